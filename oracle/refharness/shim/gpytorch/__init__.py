"""Empty symbols so that ``preds/gps.py`` (out of scope, imported by ``preds/laplace.py:11``) loads."""
