"""Runtime type checking, as in the reference (`imagescry/typechecking.py:5`)."""

from beartype import BeartypeConf, beartype

typechecker = beartype(conf=BeartypeConf(is_pep484_tower=True))
