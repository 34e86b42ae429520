"""landing_controller_b200 -- B200-native batched solver for the SRB landing NLP hot path of
se-hwan/landing-controller.  See DESIGN.md; the C ABI is in include/landing_b200.h."""
from .api import (AOS, DEVICE, DROPIN_PATH, HOST, LIB_PATH, SOA, STATUS, LandingSolver, MultiGpuSolver,  # noqa: F401
                  contact_set, dims_for, load_library, sparsity_for)
from .sweeps import (SWEEP_DT, SWEEP_N, apply_ccc_parameters, apply_sweep_parameters, grid_sweep,  # noqa: F401
                     random_sweep, single_drop, sweep_initial_guess)
from .sharding import (gather_records, pack_records, shard_bounds, shard_indices, unpack_records,  # noqa: F401
                       unshard_order)
from .schedule import (SCHED_KIN_BOX, SCHED_QF, SCHED_QX, apply_schedule_parameters, ballistic_schedule,  # noqa: F401
                       reference_schedule)
from . import sweep_io  # noqa: F401,E402
