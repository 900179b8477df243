"""Empty stand-in for the module the reference imports but does not ship (train.py:11)."""
