"""B200-native EINCM contrast-correlation objective (value + gradient).

Import as ``eincm_b200`` (see ``eincm_b200/__init__.py``).  Sub-modules:

* ``losses``  - host mirror of the reference's ``eincm.losses`` call signatures, backed by the CUDA plan
* ``plan``    - ctypes binding of the C-ABI in ``include/eincm.h``
* ``synth``   - seeded synthetic event windows (test / benchmark inputs)
"""
__version__ = '0.1.0'
