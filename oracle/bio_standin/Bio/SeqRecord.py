"""SeqRecord container (the attributes/methods InSilicoSeq touches)."""
from Bio.Seq import Seq, MutableSeq, _SeqBase


class _RestrictedDict(dict):
    """Per-letter annotations must match the sequence length (Biopython behaviour)."""

    def __init__(self, length):
        dict.__init__(self)
        self._length = int(length)

    def __setitem__(self, key, value):
        if not hasattr(value, "__len__") or not hasattr(value, "__getitem__") or (
            hasattr(self, "_length") and len(value) != self._length
        ):
            raise TypeError(
                "We only allow python sequences (lists, tuples or strings) of length %i." % self._length
            )
        dict.__setitem__(self, key, value)


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>",
                 dbxrefs=None, features=None, annotations=None, letter_annotations=None):
        self._seq = seq
        self.id = id
        self.name = name
        self.description = description
        self.dbxrefs = dbxrefs if dbxrefs is not None else []
        self.features = features if features is not None else []
        self.annotations = annotations if annotations is not None else {}
        self._per_letter_annotations = _RestrictedDict(len(seq) if seq is not None else 0)
        if letter_annotations:
            for k, v in letter_annotations.items():
                self._per_letter_annotations[k] = v

    @property
    def letter_annotations(self):
        return self._per_letter_annotations

    @letter_annotations.setter
    def letter_annotations(self, value):
        if not isinstance(value, dict):
            raise TypeError("The per-letter-annotations should be a (restricted) dictionary.")
        self._per_letter_annotations = _RestrictedDict(len(self._seq))
        for k, v in value.items():
            self._per_letter_annotations[k] = v

    @property
    def seq(self):
        return self._seq

    @seq.setter
    def seq(self, value):
        if self._per_letter_annotations:
            if len(self) != len(value):
                raise ValueError("You must empty the letter annotations first!")
            self._seq = value
        else:
            self._seq = value
            self._per_letter_annotations = _RestrictedDict(len(value))

    def __len__(self):
        return len(self._seq)

    def __iter__(self):
        return iter(self._seq)

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._seq[index]
        if isinstance(index, slice):
            rec = SeqRecord(self._seq[index], id=self.id, name=self.name, description=self.description)
            for k, v in self._per_letter_annotations.items():
                rec._per_letter_annotations[k] = v[index]
            return rec
        raise ValueError("Invalid index")

    def __str__(self):
        return "ID: %s\nDescription: %s\n%r" % (self.id, self.description, self._seq)

    def __repr__(self):
        return "SeqRecord(seq=%r, id=%r, description=%r)" % (self._seq, self.id, self.description)
