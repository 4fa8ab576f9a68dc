"""Host-side stand-in for the parts of Gridap / GridapGmsh that nuPGCM calls at set-up time."""
from .mshio import RawMesh, read_msh
from .refine import bowl_projection, refine
from .boxmesh import box_mesh
from .fem import (CellIntegrator, DiscreteModel, FacetIntegrator, LagrangeSpace, restrict,
                  restrict_vector)

__all__ = ["RawMesh", "read_msh", "refine", "bowl_projection", "box_mesh", "CellIntegrator", "DiscreteModel", "FacetIntegrator",
           "LagrangeSpace", "restrict", "restrict_vector"]
