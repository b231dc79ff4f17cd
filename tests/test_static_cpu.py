"""No linter ships in the image; this keeps the cheapest class of bug (a name a function reads but nothing binds) out of the GPU runs."""
from pathlib import Path

from tools.undefined_names import check

ROOT = Path(__file__).resolve().parents[1]


def test_no_undefined_global_names():
    files = [*ROOT.glob("uzu_b200/*.py"), *ROOT.glob("tests/*.py"), *ROOT.glob("tools/*.py"), *ROOT.glob("oracle/*.py"), ROOT / "bench.py",
             ROOT / "__graft_entry__.py"]
    bad = [(str(f.relative_to(ROOT)), *b) for f in files for b in check(str(f))]
    assert not bad, bad
