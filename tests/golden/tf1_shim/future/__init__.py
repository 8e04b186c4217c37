"""Stub of the `future` package (py2/3 compatibility imports used by the reference)."""
