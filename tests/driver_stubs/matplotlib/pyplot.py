from unittest.mock import MagicMock

_mock = MagicMock()


def __getattr__(name):
    return getattr(_mock, name)
