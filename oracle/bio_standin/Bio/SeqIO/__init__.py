"""SeqIO.parse (fasta, fastq) and SeqIO.write (fasta, fastq-sanger) with Biopython 1.81 text formats."""
from Bio.Seq import Seq
from Bio.SeqRecord import SeqRecord


def _open(handle, mode):
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        return open(handle, mode), True
    return handle, False


def _parse_fasta(fh):
    title = None
    chunks = []
    for line in fh:
        if line.startswith(">"):
            if title is not None:
                yield title, "".join(chunks)
            title = line[1:].rstrip()
            chunks = []
        elif title is not None:
            chunks.append("".join(line.split()))
        # lines before the first '>' are ignored
    if title is not None:
        yield title, "".join(chunks)


def _parse_fastq(fh):
    while True:
        head = fh.readline()
        if not head:
            return
        seq = fh.readline().rstrip("\n")
        fh.readline()
        qual = fh.readline().rstrip("\n")
        yield head[1:].rstrip(), seq, qual


def parse(handle, format):
    fh, close = _open(handle, "r")
    try:
        if format == "fasta":
            for title, seq in _parse_fasta(fh):
                try:
                    first = title.split(None, 1)[0]
                except IndexError:
                    first = ""
                yield SeqRecord(Seq(seq), id=first, name=first, description=title)
        elif format in ("fastq", "fastq-sanger"):
            for title, seq, qual in _parse_fastq(fh):
                first = title.split(None, 1)[0] if title else ""
                rec = SeqRecord(Seq(seq), id=first, name=first, description=title)
                rec.letter_annotations["phred_quality"] = [ord(c) - 33 for c in qual]
                yield rec
        else:
            raise ValueError("Unknown format '%s'" % format)
    finally:
        if close:
            fh.close()


def _clean(text):
    return text.replace("\n", " ").replace("\r", " ")


def _as_fasta(record):
    id_ = _clean(record.id)
    description = _clean(record.description)
    if description and description.split(None, 1)[0] == id_:
        title = description
    elif description:
        title = "%s %s" % (id_, description)
    else:
        title = id_
    data = str(record.seq)
    lines = [">%s\n" % title]
    for i in range(0, len(data), 60):
        lines.append(data[i:i + 60] + "\n")
    return "".join(lines)


def _as_fastq(record):
    seq_str = str(record.seq)
    quals = record.letter_annotations["phred_quality"]  # KeyError if absent
    if len(quals) != len(seq_str):
        raise ValueError("Record %s has sequence length %i but %i quality scores"
                         % (record.id, len(seq_str), len(quals)))
    if None in quals:
        raise TypeError("A quality value of None was found")
    qual_str = "".join(chr(min(93, max(0, int(round(q)))) + 33) for q in quals)
    id_ = _clean(record.id) if record.id else ""
    description = _clean(record.description)
    if description and description.split(None, 1)[0] == id_:
        title = description
    elif description:
        title = "%s %s" % (id_, description)
    else:
        title = id_
    return "@%s\n%s\n+\n%s\n" % (title, seq_str, qual_str)


def write(sequences, handle, format):
    if isinstance(sequences, SeqRecord):
        sequences = [sequences]
    fh, close = _open(handle, "w")
    count = 0
    try:
        for rec in sequences:
            if format == "fasta":
                fh.write(_as_fasta(rec))
            elif format in ("fastq", "fastq-sanger"):
                fh.write(_as_fastq(rec))
            else:
                raise ValueError("Unknown format '%s'" % format)
            count += 1
    finally:
        if close:
            fh.close()
    return count
