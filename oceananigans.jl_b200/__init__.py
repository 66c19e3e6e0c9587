"""ocean_b200 -- B200-native (sm_100a) NonhydrostaticModel time step of Oceananigans.jl on a RectilinearGrid.

The product is `libocean_b200.so` (C ABI in include/ocean_b200.h; CUDA kernels in csrc/).  This package is the
host-side mirror of the reference's user API for the hot path -- the same names, argument meaning and error
behaviour as Oceananigans (`RectilinearGrid`, `NonhydrostaticModel`, `time_step!` -> `time_step`, `set!` -> `set`,
`Simulation`, `run!` -> `run`) -- so that tests and benchmarks read like the reference's own.  The Julia extension
that makes the *real* Oceananigans dispatch onto the same C ABI is julia/OceananigansB200Ext.jl.

The directory name `oceananigans.jl_b200` is not a Python identifier; import it through the repo-root shim:
`import ocean_b200`.
"""
from ._abi import OceanB200Error, lib  # noqa: F401
from .grids import B200, CPU, RectilinearGrid, Periodic, Bounded, Flat, Float32, Float64  # noqa: F401
from .fields import (Field, FieldBoundaryConditions, FluxBoundaryCondition, ValueBoundaryCondition,  # noqa: F401
                     GradientBoundaryCondition, OpenBoundaryCondition)
from .models import (NonhydrostaticModel, Centered, WENO, ScalarDiffusivity, Smagorinsky, SmagorinskyLilly,  # noqa: F401
                     DynamicSmagorinsky, LagrangianAveraging, AnisotropicMinimumDissipation, ExplicitTimeDiscretization,
                     VerticallyImplicitTimeDiscretization, BuoyancyTracer, SeawaterBuoyancy, LinearEquationOfState, FPlane,
                     time_step, set, update_state)
from .simulations import (Simulation, run, Callback, IterationInterval, TimeInterval, TimeStepWizard,  # noqa: F401
                          conjure_time_step_wizard, NaNChecker, NPZOutputWriter, Checkpointer)
from .solvers import FFTBasedPoissonSolver, FourierTridiagonalPoissonSolver, BatchedTridiagonalSolver, solve  # noqa: F401
from .distributed import Distributed, partition_x, neighbors, gather_x, all_reduce_scalar  # noqa: F401
from .streaming import HostStreamedStepper, HostMember  # noqa: F401
from .checkpoint import checkpoint, restore  # noqa: F401
