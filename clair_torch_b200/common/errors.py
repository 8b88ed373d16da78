"""Error types of the public mirror.

Every public function of the reference is `@typechecked` (typeguard 4): a wrong argument type raises
`typeguard.TypeCheckError`, which is NOT a `TypeError` subclass.  `ArgumentTypeError` derives from both, so
`except typeguard.TypeCheckError` written against the reference and `except TypeError` written against this package
catch the same mistakes.  Without typeguard installed it is a plain `TypeError`.
"""
try:
    from typeguard import TypeCheckError as _TypeCheckError

    class ArgumentTypeError(_TypeCheckError, TypeError):
        pass
except ImportError:                                                  # pragma: no cover - typeguard ships with the image

    class ArgumentTypeError(TypeError):
        pass
