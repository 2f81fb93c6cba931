"""Minimal stand-in for the parts of the reference's `dnnlib` the hot path touches (S3/dnnlib/util.py: EasyDict)."""


class EasyDict(dict):
    """dict with attribute access."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]
