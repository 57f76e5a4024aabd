"""The reference resolves every component through a plain dict plus a `build_*` function that takes the `type` entry
out of a config sub-dict and splats the rest into the constructor.  The tables stay plain dicts
(install_into_dmb() merges them into the reference's); the look-up itself lives here once."""


def lookup(table, what, name):
    """The class registered under `name`; unknown names fail with an AssertionError that lists the table's keys, like
    the reference's builders (e.g. disp_predictors/builder.py:16-17)."""
    if name not in table:
        raise AssertionError("%s type not found, expected one of %s, but got %r" % (what, sorted(table), name))
    return table[name]


def ctor_kwargs(section, **extra):
    """Constructor keyword arguments of a config section: everything but its `type`, plus `extra`."""
    kwargs = {key: section[key] for key in section if key != "type"}
    kwargs.update(extra)
    return kwargs
