"""Drop-in shim: `from ssqueeze import _rs` (src/ssqueeze/__init__.py:1-27 of the
reference) resolves to the B200 engine.  Unlike the reference shim there is no
dummy fallback: a missing/unbuilt library raises ImportError."""
from ssqueeze_rs_b200 import _rs  # noqa: F401

__all__ = ["_rs"]


def main():  # console script `ssqueeze:main` (pyproject.toml:20-21)
    print(_rs.hello_from_bin())
