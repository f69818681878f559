"""``PySolvers.Linear`` surface (PySolvers/Linear/__init__.py:1-12)."""
from .base import (IterativeLinearSolver, IterativeLinearSolverType,  # noqa: F401
                   LinearSolver, LinearSolverType, mvmult)
from .precond import (Preconditioner, GenericPreconditioner,  # noqa: F401
                      LeftPreconditioner, RightPreconditioner,
                      IdentityPreconditioner, PreconditionerType,
                      IdentityPreconditionerType)
from .krylov import PCG, PCGSolver, GMRES, GMRESSolver  # noqa: F401
from .factorized import (RightIC, ICRightPreconditioner, LeftILUT, RightILUT,  # noqa: F401
                         ILUTPreconditioner, LeftILUTPreconditioner,
                         RightILUTPreconditioner)
from .multigrid import (AMG, AMGPreconditioner, AMGVCycle, AMGVCycleSolver,  # noqa: F401
                        JacobiSmoother, DampedJacobiSmoother, GaussSeidelSmoother,
                        MLHierarchy, SmoothedAggregationMLHierarchy, VCycleManager,
                        SA_coarsen, makeRestrictionOp)
from .direct import DefaultDirect, DefaultDirectSolver  # noqa: F401
