"""Stand-in for matplotlib (not installed in the image): every attribute is a permissive mock, nothing is drawn."""
from unittest.mock import MagicMock

_mock = MagicMock()


def use(*args, **kwargs):
    return None


def __getattr__(name):
    return getattr(_mock, name)
