"""Game-spec mini language ``key_val:key_val`` (reference: rlpoker/games/util.py:4-26)."""


def parse_params(spec):
    params = {}
    for part in filter(None, spec.split(':')):
        key, value = part.split('_', 1)
        params[key] = value
    return params


class ExtensiveGameBuilder:
    @staticmethod
    def build(spec):
        raise NotImplementedError()
