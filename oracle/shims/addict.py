"""Ten-line stand-in for the `addict` package (imported by the reference's
utils/config_handler.py; absent in this image).  TEST INFRASTRUCTURE."""


class Dict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
