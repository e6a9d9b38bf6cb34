"""swmm_b200 -- B200-native dynamic-wave flow routing + water-quality transport for SWMM 5.2.4.

Only the hot path lives here: csrc/ holds the sm_100a kernels and the C-ABI,
seam/ the shim that plugs them behind the reference's dynwave_* / qualrout_* seam, and the Python
modules are the thin host side (ctypes over the C-ABI, scenario generators, network builder).
"""
__version__ = "0.1.0"
