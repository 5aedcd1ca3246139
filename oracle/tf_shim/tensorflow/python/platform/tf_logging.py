"""`tensorflow.python.platform.tf_logging` stand-in for the oracle's numpy TF shim (test infrastructure)."""


def debug(*args, **kwargs):
    pass


info = warning = error = debug
