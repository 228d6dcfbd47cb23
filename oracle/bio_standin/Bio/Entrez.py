"""Import-only placeholder (iss/download.py imports Bio.Entrez; network is out of scope)."""
email = None
tool = "standin"


def _unavailable(*a, **k):
    raise RuntimeError("Bio.Entrez is not available in the offline stand-in")


esearch = efetch = read = _unavailable
