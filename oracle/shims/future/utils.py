"""`future.utils.with_metaclass` stand-in (reference graphtools/base.py:5)."""


def with_metaclass(meta, *bases):
    class _Tmp(meta):
        def __new__(mcls, name, this_bases, d):
            return meta(name, bases, d)

        @classmethod
        def __prepare__(mcls, name, this_bases):
            return meta.__prepare__(name, bases)

    return type.__new__(_Tmp, "temporary_class", (), {})
