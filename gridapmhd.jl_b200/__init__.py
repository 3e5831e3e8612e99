"""mhd-b200: B200-native assembly + Krylov kernels for the GridapMHD H1-HDiv hot path.

Directory name follows the build contract (`gridapmhd.jl_b200/`); since a dot is not importable the
package registers itself as `gridapmhd_jl_b200` (see the root shim `gridapmhd_jl_b200.py`).
"""
__version__ = "0.1.0"
