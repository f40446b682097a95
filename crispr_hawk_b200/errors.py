"""Error convention of the reference (exception_handlers.py:28-58, crisprhawk_error.py):
`debug=True` raises `exc_type("\\n\\n" + message)`, otherwise the message goes to
stderr in red and the process exits with an `os.EX_*` code. When the reference
package is importable its own handler and exception classes are used, so callers
catching `CrisprHawkIupacTableError` keep working."""

from __future__ import annotations

import sys
from typing import NoReturn, Optional


class CrisprHawkError(Exception):
    pass


class CrisprHawkIupacTableError(CrisprHawkError):
    pass


class CrisprHawkPamError(CrisprHawkError):
    pass


class CrisprHawkCfdScoreError(CrisprHawkError):
    pass


class CrisprHawkGuideError(CrisprHawkError):
    pass


class CrisprHawkGcContentError(CrisprHawkError):
    pass


class CrisprHawkAnnotationError(CrisprHawkError):
    pass


def _reference_module(name: str):
    try:
        import importlib

        return importlib.import_module(f"crisprhawk.{name}")
    except Exception:
        return None


def error_class(name: str):
    mod = _reference_module("crisprhawk_error")
    if mod is not None and hasattr(mod, name):
        return getattr(mod, name)
    return globals()[name]


def exception_handler(
    exception_type: type, exception: str, code: int, debug: bool, e: Optional[Exception] = None
) -> NoReturn:
    mod = _reference_module("exception_handlers")
    if mod is not None:
        mod.exception_handler(exception_type, exception, code, debug, e)
    if debug:
        if e:
            raise exception_type(f"\n\n{exception}") from e
        raise exception_type(f"\n\n{exception}")
    sys.stderr.write(f"\033[31m\n\nERROR: {exception}\n\033[39m")
    sys.exit(code)


def print_verbosity(message: str, verbosity: int, threshold: int) -> None:
    """utils.py:165-182"""
    if verbosity >= threshold:
        sys.stdout.write(f"{message}\n")
