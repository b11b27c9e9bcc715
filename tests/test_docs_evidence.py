"""Every evidence file the documents cite under profiles/ exists (names with {a,b} alternatives or * wildcards are expanded)."""
import glob
import itertools
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expand(name):
    parts = re.split(r"\{([^}]*)\}", name)
    if len(parts) == 1:
        return [name]
    alts = [p.split(",") if i % 2 else [p] for i, p in enumerate(parts)]
    return ["".join(c) for c in itertools.product(*alts)]


def test_cited_profiles_exist():
    missing = []
    for doc in ("DESIGN.md", "README.md", "profiles/README.md", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, doc)).read()
        cited = set(re.findall(r"`(?:profiles/)?(r0\d[a-z0-9]*_[A-Za-z0-9_{},.*]+)`", text))
        for name in cited:
            for n in _expand(name):
                if n.endswith("_"):
                    n += "*"
                if not glob.glob(os.path.join(ROOT, "profiles", n)):
                    missing.append((doc, n))
    assert not missing, missing
