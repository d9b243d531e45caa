"""Stand-in for the third-party ``geographiclib`` package, which the reference imports at module
top (``utils.py:4``) and which is not installed in this image.  TEST / BENCH INFRASTRUCTURE ONLY:
it lets ``oracle/reference_arm.py`` import the unmodified reference from ``baseline/_ref``; every
track that goes through it is built from arrays, so nothing here is ever called."""
