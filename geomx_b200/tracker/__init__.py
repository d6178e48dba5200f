"""Job launchers (``python -m geomx_b200.tracker.launch``): build the ``DMLC_*`` environment of every HiPS role and start the processes.

Parity: ``3rdparty/ps-lite/tracker/{dmlc_local,dmlc_ssh,dmlc_mpi}.py`` + ``tracker.py`` (generic dmlc launchers; the reference's own demo
scripts hand-write the same environments, ``scripts/gpu/run_vanilla_hips.sh:8-148``)."""
from .launch import HipsJob, launch  # noqa: F401
