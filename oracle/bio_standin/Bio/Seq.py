"""Seq / MutableSeq containers (container semantics of Biopython 1.81)."""


class _SeqBase:
    __slots__ = ("_data",)

    def __len__(self):
        return len(self._data)

    def __str__(self):
        return "".join(self._data) if isinstance(self._data, list) else self._data

    def __repr__(self):
        return "%s(%r)" % (type(self).__name__, str(self))

    def __iter__(self):
        return iter(self._data)

    def __eq__(self, other):
        if isinstance(other, _SeqBase):
            return str(self) == str(other)
        if isinstance(other, str):
            return str(self) == other
        return NotImplemented

    def __hash__(self):
        return hash(str(self))

    def __contains__(self, item):
        return str(item) in str(self)

    def __add__(self, other):
        return type(self)(str(self) + str(other))

    def __radd__(self, other):
        return type(self)(str(other) + str(self))

    def startswith(self, prefix, *a):
        return str(self).startswith(str(prefix), *a)

    def endswith(self, suffix, *a):
        return str(self).endswith(str(suffix), *a)

    def upper(self):
        return type(self)(str(self).upper())

    def lower(self):
        return type(self)(str(self).lower())

    def count(self, sub):
        return str(self).count(str(sub))


class Seq(_SeqBase):
    """Immutable sequence; indexing with an int gives a 1-char str, a slice gives a Seq."""

    def __init__(self, data=""):
        if isinstance(data, _SeqBase):
            data = str(data)
        elif isinstance(data, (bytes, bytearray)):
            data = data.decode("ascii")
        elif not isinstance(data, str):
            raise TypeError("data should be a string, bytes, bytearray, Seq, or MutableSeq object")
        self._data = data

    def __getitem__(self, index):
        if isinstance(index, slice):
            return Seq(self._data[index])
        return self._data[index]


class MutableSeq(_SeqBase):
    """Mutable sequence backed by a list of 1-char strings."""

    def __init__(self, data=""):
        if isinstance(data, _SeqBase):
            data = str(data)
        elif isinstance(data, (bytes, bytearray)):
            data = data.decode("ascii")
        elif not isinstance(data, str):
            raise TypeError("data should be a string, bytes, bytearray, Seq, or MutableSeq object")
        self._data = list(data)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return MutableSeq("".join(self._data[index]))
        return self._data[index]  # IndexError when out of range, like Biopython

    def __setitem__(self, index, value):
        if isinstance(index, slice):
            self._data[index] = list(str(value))
        else:
            value = str(value)
            if len(value) != 1:
                raise ValueError("expected a single letter")
            self._data[index] = value

    def __delitem__(self, index):
        del self._data[index]

    def append(self, c):
        self._data.append(str(c))

    def insert(self, i, c):
        self._data.insert(i, str(c))

    def pop(self, i=-1):
        return self._data.pop(i)

    def extend(self, other):
        self._data.extend(list(str(other)))
