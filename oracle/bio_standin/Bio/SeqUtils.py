"""gc_fraction as in Biopython 1.81 (default ambiguous="remove"): returns a 0-1 fraction."""


def gc_fraction(seq, ambiguous="remove"):
    s = str(seq)
    if ambiguous not in ("remove", "ignore", "weighted"):
        raise ValueError("ambiguous value '%s' not recognized" % ambiguous)
    gc = sum(s.count(x) for x in "CGScgs")
    if ambiguous == "remove":
        length = gc + sum(s.count(x) for x in "ATWUatwu")
    else:
        length = len(s)
    if length == 0:
        return 0
    return gc / length
