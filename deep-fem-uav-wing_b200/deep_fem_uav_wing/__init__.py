"""deep_fem_uav_wing - B200-native drop-in for the GraphSAGE hot path of Deep-FEM-UAV-Wing.

Only the ``gnn`` sub-package exists here: geometry, meshing and FEM stay with the reference
(external Blender / Gmsh / CalculiX binaries) and are out of scope.
"""
__all__ = ["__version__"]
__version__ = "0.1.0"
