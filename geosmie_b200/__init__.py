"""geosmie_b200 -- B200-native (sm_100a CUDA, FP64) implementation of the GEOSmie Mie lookup-table hot path.

The compute lives in ``libgeosmie_b200.so`` (hand-written CUDA behind the C ABI of ``include/geosmie_b200.h``) and is
reached through ctypes (``geosmie_b200._lib``).  There is no CPU fallback: importing the host-side mirrors works
anywhere, any compute call raises if the library or a B200 is missing.

Host-side mirrors of the reference interface (same names, arguments and error behaviour):
  geosmie_b200.pymiecoated     Mie, mie_coated.MultipleMie            (src/pymiecoated/pymiecoated)
  geosmie_b200.dointegration   fun, rawMie, integratePSD, ...         (src/geosmie/dointegration.py)
  geosmie_b200.particleparams / hydrophobic / bandaverage / runoptics / runbands   (src/geosmie/*.py)
  geosmie_b200.gsf.convertncdf / rungsf                                (src/gsf/*.py)
"""
__version__ = "0.1.0"
