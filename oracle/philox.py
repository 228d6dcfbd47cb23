"""Philox4x32-10 (Salmon et al., SC'11; same constants as Random123 / cuRAND) and the
draw-slot schedule ("RNG contract v2") shared by every implementation in this repo.

TEST INFRASTRUCTURE ONLY (oracle/).  The product computes the same function on the GPU
(insilicoseq_b200/csrc/philox.cuh); nothing in the product imports this file.

RNG contract v2 (v1 gave every indel site a slot: INS block = position, DEL word = position; v2 ranks the non-zero sites)
---------------
The reference draws from two *sequential* MT19937 streams (stdlib `random`, legacy
`numpy.random`) seeded with `seed + cpu_number` per worker (iss/generator.py:234-236).
A GPU needs position-addressed draws, so every draw site of the hot path is given a
fixed slot in a counter-based Philox4x32-10 stream:

    key     = (K & 0xffffffff, K >> 32)         K = (seed + cpu_number) mod 2**64
    counter = (block, stream | attempt << 8, P & 0xffffffff, P >> 32)
              P = ordinal of the pair inside its worker (sum of n_pairs of the preceding
                  work items of that worker + i)
    one Philox call yields 4 words w0..w3.

    stream 0 (PAIR)   block 0: w0 insert size       (kde.py:97  np.random.rand)
                               w1 quality bin, forward (kde.py:74 np.random.choice)
                               w2 quality bin, reverse
                               w3 gc_bias accept draw (generator.py:88 np.random.rand)
                      block 1: (w0,w1) forward start (generator.py:135 random.randrange)
                               (w2,w3) fallback start (generator.py:144)
                      block 2: (w0,w1) fallback reverse end (generator.py:175)
                               (w2,w3) fragment length   (generator.py:122 np.random.normal)
    per read o (0 forward, 1 reverse), stream = 1 + 4*o + k:
      k=0 QUAL  block b : w0 phred(2b)  w1 subst-test(2b)  w2 phred(2b+1)  w3 subst-test(2b+1)
                          (kde.py:84 np.random.rand ; error_models/__init__.py:94 random.random)
      k=1 SUBC  block b : w_j substitution choice at position 4b+j   (__init__.py:95-97)
      k=2 DEL   rank r -> word r&3 of block r>>2; r = rank of the loop position among the positions 0..L-2
                          at which ANY base has a deletion probability > 0            (__init__.py:209)
      k=3 INS   rank r -> word r&3 of block r>>2; r = rank of (loop position, j-th key of the insertion dict)
                          among the entries with probability > 0, in (position, j) order  (__init__.py:193-194)
                A site whose probability is not > 0 has no slot: its test `random() < p` is false whatever is drawn.

    uniform()         := w * 2**-32           (exact in f64)
    randrange(a, b)   := a + ((w_lo | w_hi << 32) * (b - a) >> 64); ValueError if b <= a
    `attempt` counts re-simulations of the same pair under --gc_bias (generator.py:82-92).

Slots are position-addressed, so a slot whose thresholds are all zero may be skipped
without changing any result.
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF

# stream ids
S_PAIR = 0
K_QUAL, K_SUBC, K_DEL, K_INS = 0, 1, 2, 3


def read_stream(orientation, kind):
    """orientation 0/1 (forward/reverse), kind one of K_*."""
    return 1 + 4 * orientation + kind


def philox4x32_10(counter, key):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (python ints)."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return (c0, c1, c2, c3)


def make_key(seed, cpu_number):
    k = (int(seed) + int(cpu_number)) & 0xFFFFFFFFFFFFFFFF
    return (k & MASK32, k >> 32)


def block(key, pair_ordinal, stream, blk, attempt=0):
    ctr = (blk & MASK32, (stream | (attempt << 8)) & MASK32, pair_ordinal & MASK32, (pair_ordinal >> 32) & MASK32)
    return philox4x32_10(ctr, key)


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised variant over numpy uint64 arrays holding 32-bit values (for bulk tests)."""
    import numpy as np

    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 = np.asarray(k0, dtype=np.uint64)
    k1 = np.asarray(k1, dtype=np.uint64)
    m = np.uint64(MASK32)
    s = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = ((p1 >> s) ^ c1 ^ k0) & m
        n1 = p1 & m
        n2 = ((p0 >> s) ^ c3 ^ k1) & m
        n3 = p0 & m
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3
