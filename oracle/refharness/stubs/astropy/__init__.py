"""Minimal stand-in for the `astropy` package (TEST INFRASTRUCTURE ONLY).

The Lightspinner reference imports `astropy.units` for unit tagging in its
*setup* code (atmosphere.py, fal.py, utils.py).  astropy is not installed in
this image and there is no network, so `oracle/refharness` puts this stub on
`sys.path` in order to import the unmodified reference from /root/reference
when golden vectors are generated.  Nothing in the product path imports it.
"""
from . import units  # noqa: F401
