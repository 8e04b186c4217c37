from collections import OrderedDict  # noqa: F401
