"""Host prototype of the NEXT compressor step (not used by the product): one gzip member = one dynamic-Huffman block with
literals plus at most one LZ77 match per FASTQ record -- the name line against the previous record's name line -- and
optional distance-1 runs.  Fixed-size header like csrc/iss_deflate.cu (flat 4-bit code-length code), HLIT = 29, HDIST = 29:
17 + 19*3 + (286 + 30)*4 = 1338 header bits.  Checks itself by decompressing with zlib.

    python tools/gz_lz_prototype.py tests/golden/miseq150_bench_model/R1.fastq
"""
import heapq
import sys
import zlib

LB = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LE = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DB = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
      12289, 16385, 24577]
DE = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def bucket(v, base):
    k = len(base) - 1
    while v < base[k]:
        k -= 1
    return k


def limited_lengths(freq, maxbits=15):
    """Huffman code lengths, limited to maxbits the zlib way (gen_bitlen): cap while descending, repair the counts,
    re-deal the lengths to the symbols in order of increasing frequency.  Every symbol with freq > 0 gets a code."""
    syms = sorted((f, s) for s, f in enumerate(freq) if f)
    n = len(syms)
    lens = [0] * len(freq)
    if n == 1:
        lens[syms[0][1]] = 1
        return lens
    heap = [(f, i, None, None) for i, (f, _) in enumerate(syms)]
    heapq.heapify(heap)
    nxt = n
    parent = {}
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        parent[a[1]] = nxt
        parent[b[1]] = nxt
        heapq.heappush(heap, (a[0] + b[0], nxt, None, None))
        nxt += 1
    root = nxt - 1
    depth = {root: 0}
    overflow = 0
    bl = [0] * (maxbits + 2)
    for node in range(root - 1, -1, -1):
        d = depth[parent[node]] + 1
        if d > maxbits:
            d = maxbits
            overflow += 1
        depth[node] = d
        if node < n:
            bl[d] += 1
    if overflow:
        while overflow > 0:
            bits = maxbits - 1
            while bl[bits] == 0:
                bits -= 1
            bl[bits] -= 1
            bl[bits + 1] += 2
            bl[maxbits] -= 1
            overflow -= 2
        i = 0
        for bits in range(maxbits, 0, -1):
            for _ in range(bl[bits]):
                lens[syms[i][1]] = bits
                i += 1
    else:
        for i in range(n):
            lens[syms[i][1]] = depth[i]
    return lens


def canonical(lens):
    bl = [0] * 17
    for l in lens:
        bl[l] += 1
    bl[0] = 0
    code = 0
    first = [0] * 17
    for b in range(1, 16):
        code = (code + bl[b - 1]) << 1
        first[b] = code
    out = []
    for l in lens:
        if l:
            out.append(int(format(first[l], "0%db" % l)[::-1], 2))  # stored bit-reversed: DEFLATE packs codes MSB first
            first[l] += 1
        else:
            out.append(0)
    return out


class Bits(object):
    def __init__(self):
        self.acc = 0
        self.n = 0

    def put(self, v, nbits):
        self.acc |= v << self.n
        self.n += nbits

    def bytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def tokens_fastq(data, runs):
    """('L', byte) / ('M', length, distance): the name line of a record against the previous record's, optionally runs."""
    starts, p = [], 0
    for li, line in enumerate(data.split(b"\n")[:-1]):
        if li % 4 == 0:
            starts.append(p)
        p += len(line) + 1
    prev = {a: b for a, b in zip(starts[1:], starts[:-1])}
    out, i, n = [], 0, len(data)
    while i < n:
        m = d = 0
        if i in prev:
            ps = prev[i]
            while i + m < n and m < 258 and data[i + m] == data[ps + m] and data[i + m] != 10:
                m += 1
            d = i - ps
        if m < 3 and runs and i > 0:
            r = 0
            while i + r < n and r < 258 and data[i + r] == data[i - 1]:
                r += 1
            if r >= 4:
                m, d = r, 1
        if m >= 3:
            out.append(("M", m, d))
            i += m
        else:
            out.append(("L", data[i]))
            i += 1
    return out


def member(data, runs=False):
    toks = tokens_fastq(data, runs)
    lf = [0] * 286
    df = [0] * 30
    for t in toks:
        if t[0] == "L":
            lf[t[1]] += 1
        else:
            lf[257 + bucket(t[1], LB)] += 1
            df[bucket(t[2], DB)] += 1
    lf[256] = 1
    if not any(df):
        df[0] = 1  # a block needs one distance code even without matches
    ll, dl = limited_lengths(lf), limited_lengths(df)
    lc, dc = canonical(ll), canonical(dl)
    b = Bits()
    b.put(1, 1); b.put(2, 2); b.put(29, 5); b.put(29, 5); b.put(15, 4)  # BFINAL, BTYPE=10, HLIT=29, HDIST=29, HCLEN=15
    for k in range(19):  # code-length-code lengths in the order 16,17,18,0,8,7,...: three zeros, sixteen fours
        b.put(0 if k < 3 else 4, 3)
    for l in ll + dl:  # each length as the 4-bit code of its value (code == value, MSB first)
        b.put(int(format(l, "04b")[::-1], 2), 4)
    assert b.n == 17 + 19 * 3 + (286 + 30) * 4
    for t in toks:
        if t[0] == "L":
            b.put(lc[t[1]], ll[t[1]])
        else:
            k = bucket(t[1], LB)
            b.put(lc[257 + k], ll[257 + k]); b.put(t[1] - LB[k], LE[k])
            k = bucket(t[2], DB)
            b.put(dc[k], dl[k]); b.put(t[2] - DB[k], DE[k])
    b.put(lc[256], ll[256])
    body = b.bytes()
    crc = zlib.crc32(data) & 0xFFFFFFFF
    return bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF]) + body + crc.to_bytes(4, "little") + (len(data) & 0xFFFFFFFF).to_bytes(4, "little")


if __name__ == "__main__":
    data = open(sys.argv[1], "rb").read()
    chunk = 65536
    for runs in (False, True):
        z = b"".join(member(data[o:o + chunk], runs) for o in range(0, len(data), chunk))
        assert zlib.decompress(z, 16 + zlib.MAX_WBITS) == data[:chunk] and __import__("gzip").decompress(z) == data
        print("runs=%s: %d -> %d bytes, ratio %.2f (zlib -6: %.2f)" % (runs, len(data), len(z), len(data) / len(z),
                                                                       len(data) / len(zlib.compress(data, 6))))
