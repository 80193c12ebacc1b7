"""Ground-truth landscapes that are pure tables, evaluated on the GPU.

``flexs.landscapes.{TFBinding, AdditiveAAVPackaging}`` of the reference (flexs/landscapes/__init__.py:3,7)
with the same constructors, ``registry()`` functions and ``get_fitness`` results.  The reference's other
landscapes (ViennaRNA, PyRosetta, TAPE) wrap third-party simulators and are out of scope.
The measurement files are the reference's data, not part of this package: pass the file, or point
``FLEXS_DATA_DIR`` at a checkout's ``flexs/landscapes/data`` directory.
"""
from flexs_b200.landscapes import additive_aav_packaging, tf_binding  # noqa: F401
from flexs_b200.landscapes.additive_aav_packaging import AdditiveAAVPackaging  # noqa: F401
from flexs_b200.landscapes.table_landscape import DeviceTableLandscape, data_dir  # noqa: F401
from flexs_b200.landscapes.tf_binding import TFBinding  # noqa: F401
