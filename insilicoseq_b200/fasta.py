"""FASTA reading/writing with the semantics the reference gets from Bio.SeqIO (iss/generator.py:270,475-476):
record.id = first whitespace-delimited token of the title line, sequence = all following lines with
whitespace removed, case preserved.  Sequences are kept as bytes (they go straight to the GPU)."""


class Record(object):
    """The slice of Bio.SeqRecord the generator touches: .id, .description, .seq (bytes here), len()."""

    __slots__ = ("id", "description", "seq")

    def __init__(self, seq, id="", description=""):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        self.seq = seq
        self.id = id
        self.description = description

    def __len__(self):
        return len(self.seq)

    def __repr__(self):
        return "Record(id=%r, len=%d)" % (self.id, len(self.seq))


_WS = b" \t\n\r\x0b\x0c"  # what bytes.split() treats as white space


class TextSeq(object):
    """The sequence of a record whose bases went from the FASTA text to the GPU without passing through Python
    (Engine.load_fasta): len() is the base count the device reported; the bytes are cut out of the (memory-mapped)
    text only if somebody asks for them."""

    __slots__ = ("_buf", "_off", "_len", "_n")

    def __init__(self, buf, off, length, n_bases):
        self._buf, self._off, self._len, self._n = buf, off, length, n_bases

    def __len__(self):
        return self._n

    def __bytes__(self):
        return bytes(self._buf[self._off:self._off + self._len]).translate(None, _WS)

    def __str__(self):
        return bytes(self).decode("ascii")

    def __eq__(self, other):
        return bytes(self) == (other.encode("ascii") if isinstance(other, str) else bytes(other))

    def __getitem__(self, key):
        return bytes(self)[key]


def parse(handle):
    """Yield Record objects from a path or an open text/binary handle.

    A path is memory-mapped and cut at line-initial '>' with `find`; a record's sequence is its body with white space
    deleted by one `bytes.translate` (C speed), so a multi-gigabase FASTA is not walked line by line in Python."""
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        import mmap

        with open(handle, "rb") as fh:
            try:
                buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except ValueError:  # empty file
                return
            try:
                for rec in _records(buf):
                    yield rec
            finally:
                buf.close()
        return
    data = handle.read()
    if isinstance(data, str):
        data = data.encode("ascii")
    for rec in _records(data):
        yield rec


def _records(buf):
    n = len(buf)
    if n == 0:
        return
    pos = 0 if buf[0:1] == b">" else buf.find(b"\n>") + 1  # text before the first title line is ignored
    if pos == 0 and buf[0:1] != b">":
        return
    while pos < n:
        nxt = buf.find(b"\n>", pos)
        end = nxt + 1 if nxt >= 0 else n
        eol = buf.find(b"\n", pos, end)
        if eol < 0:
            eol = end
        yield _make(bytes(buf[pos + 1:eol]).rstrip(), [bytes(buf[eol + 1:end]).translate(None, _WS)])
        pos = end


def _make(title, chunks):
    t = title.decode("ascii", errors="replace")
    parts = t.split(None, 1)
    rid = parts[0] if parts else ""
    return Record(b"".join(chunks), id=rid, description=t)


def write(records, handle, width=60):
    """FASTA text as Bio.SeqIO.write(..., 'fasta') lays it out (60 columns)."""
    n = 0
    for r in records:
        desc = r.description or ""
        if desc and desc.split(None, 1)[0] == r.id:
            title = desc
        elif desc:
            title = "%s %s" % (r.id, desc)
        else:
            title = r.id
        handle.write(">%s\n" % title)
        s = r.seq.decode("ascii")
        for i in range(0, len(s), width):
            handle.write(s[i:i + width] + "\n")
        n += 1
    return n
