"""Import stand-in for `omegaconf` (type annotations of the reference's head modules only).  TEST INFRASTRUCTURE ONLY."""


class DictConfig(dict):
    pass


class ListConfig(list):
    pass
