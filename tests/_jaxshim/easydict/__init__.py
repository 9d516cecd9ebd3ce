"""TEST INFRASTRUCTURE ONLY - stand-in for the `easydict` package (absent from this image) so that the reference's
src/evaluations/theta_eval.py imports: a dict whose keys read as attributes, nested dicts converted on the way in."""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, EasyDict(v) if isinstance(v, dict) and not isinstance(v, EasyDict) else v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
