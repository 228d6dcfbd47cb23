"""FASTA reading/writing with the semantics the reference gets from Bio.SeqIO (iss/generator.py:270,475-476):
record.id = first whitespace-delimited token of the title line, sequence = all following lines with
whitespace removed, case preserved.  Sequences are kept as bytes (they go straight to the GPU)."""


class Record(object):
    """The slice of Bio.SeqRecord the generator touches: .id, .description, .seq (bytes here), len()."""

    __slots__ = ("id", "description", "seq", "gpu_index")

    def __init__(self, seq, id="", description=""):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        self.seq = seq
        self.id = id
        self.description = description
        self.gpu_index = None

    def __len__(self):
        return len(self.seq)

    def __repr__(self):
        return "Record(id=%r, len=%d)" % (self.id, len(self.seq))


def parse(handle):
    """Yield Record objects from a path or an open text/binary handle."""
    close = False
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        handle = open(handle, "rb")
        close = True
    try:
        title = None
        chunks = []
        for line in handle:
            if isinstance(line, str):
                line = line.encode("ascii")
            if line.startswith(b">"):
                if title is not None:
                    yield _make(title, chunks)
                title = line[1:].rstrip()
                chunks = []
            elif title is not None:
                chunks.append(b"".join(line.split()))
        if title is not None:
            yield _make(title, chunks)
    finally:
        if close:
            handle.close()


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
