"""Environment-variable flag registry — every GeoMX / ps-lite / kvstore variable of the reference, with defaults.

Parity: ``docs/source/env-var-summary.rst:4-142`` and the ``dmlc::GetEnv`` / ``ps::Environment::find`` call sites
enumerated in SURVEY §5.6.  ``describe()`` prints the table; ``get(name)`` returns the typed value.  Flags marked
``ADAPTIVE_K_FLAG`` / ``DMLC_K_MIN`` / ``DGT_INFO`` are parsed but unused by the reference; they are given a meaning here."""
from __future__ import annotations

import os

_F = {}


def _reg(name, default, typ, doc, honoured=True):
    _F[name] = (default, typ, doc, honoured)


# roles / topology ------------------------------------------------------------------------------------------
_reg("DMLC_ROLE", "worker", str, "worker | server | scheduler")
_reg("DMLC_ROLE_GLOBAL", "", str, "global_server | global_scheduler (central party)")
_reg("DMLC_ROLE_MASTER_WORKER", 0, int, "1 on the single worker of the central party")
_reg("DMLC_PS_ROOT_URI", "127.0.0.1", str, "local scheduler address")
_reg("DMLC_PS_ROOT_PORT", 9091, int, "local scheduler port")
_reg("DMLC_PS_GLOBAL_ROOT_URI", "127.0.0.1", str, "global scheduler address")
_reg("DMLC_PS_GLOBAL_ROOT_PORT", 9092, int, "global scheduler port")
_reg("DMLC_NUM_SERVER", 1, int, "servers per party (local tier allows exactly 1; central party = #global servers)")
_reg("DMLC_NUM_WORKER", 1, int, "workers per party")
_reg("DMLC_NUM_GLOBAL_SERVER", 1, int, "global servers (MultiGPS)")
_reg("DMLC_NUM_GLOBAL_WORKER", 1, int, "parties (= local servers acting as global workers)")
_reg("DMLC_NUM_ALL_WORKER", 1, int, "training workers over all parties")
_reg("DMLC_ENABLE_CENTRAL_WORKER", 0, int, "central-party workers also train")
_reg("DMLC_INTERFACE", "", str, "network interface to bind")
_reg("DMLC_NODE_HOST", "", str, "explicit host/IP of this node")
_reg("PORT", 0, int, "explicit port of this node")
_reg("DMLC_LOCAL", 0, int, "use unix-domain sockets instead of TCP (reference: ipc://)")
_reg("DMLC_USE_KUBERNETES", 0, int, "bind 0.0.0.0")
_reg("DMLC_RANK", -1, int, "set by the Van after registration")
_reg("DMLC_GLOBAL_RANK", -1, int, "set by the Van after global registration")
# transport -------------------------------------------------------------------------------------------------
_reg("PS_VERBOSE", 0, int, "1|2: log control / all messages")
_reg("PS_RESEND", 0, int, "enable ACK + timeout resend")
_reg("PS_RESEND_TIMEOUT", 1000, int, "ms")
_reg("PS_DROP_MSG", 0, int, "percentage of received messages to drop (fault injection)")
_reg("PS_HEARTBEAT_INTERVAL", 0, int, "seconds, 0 = off")
_reg("PS_HEARTBEAT_TIMEOUT", 0, int, "seconds")
# kvstore ---------------------------------------------------------------------------------------------------
_reg("MXNET_KVSTORE_BIGARRAY_BOUND", 1000000, int, "elements; arrays ≥ bound are partitioned over all (global) servers")
_reg("MXNET_KVSTORE_SIZE_LOWER_BOUND", 200000, int, "elements; BSC/MPQ only above this size")
_reg("MXNET_KVSTORE_USE_HFA", 0, int, "hierarchical frequency aggregation")
_reg("MXNET_KVSTORE_HFA_K1", 1, int, "local steps per local sync")
_reg("MXNET_KVSTORE_HFA_K2", 1, int, "local syncs per global sync")
_reg("MXNET_KVSTORE_REDUCTION_NTHREADS", 4, int, "host reduction threads")
_reg("MXNET_KVSTORE_SERIAL_PUSH", 0, int, "serialise pushes")
_reg("MXNET_KVSTORE_USETREE", 0, int, "tree reduce (uniform on NVSwitch → flat)")
_reg("MXNET_KVSTORE_TREE_ARRAY_BOUND", 10000000, int, "", False)
_reg("MXNET_KVSTORE_TREE_BACKTRACK", 0, int, "", False)
_reg("MXNET_KVSTORE_TREE_LINK_USAGE_PENALTY", 0.7, float, "", False)
_reg("MXNET_KVSTORE_LOGTREE", 0, int, "", False)
_reg("MXNET_KVSTORE_DIST_ROW_SPARSE_VERBOSE", 0, int, "")
_reg("MXNET_ENABLE_GPU_P2P", 1, int, "enable peer access")
_reg("MXNET_UPDATE_AGGREGATION_SIZE", 16, int, "keys per grouped collective")
# accelerators ----------------------------------------------------------------------------------------------
_reg("ENABLE_P3", 0, int, "priority-based parameter propagation")
_reg("ENABLE_DGT", 0, int, "1 lossy channels | 2 prioritised reliable | 3 + 4-bit encode")
_reg("DMLC_UDP_CHANNEL_NUM", 3, int, "number of low-priority channels")
_reg("DMLC_K", 0.5, float, "fraction of blocks on the reliable channel")
_reg("DMLC_K_MIN", 0.2, float, "lower bound of the important fraction when ADAPTIVE_K_FLAG=1 (fabric DGT)", True)
_reg("ADAPTIVE_K_FLAG", 0, int, "fabric DGT: DMLC_K is a share of the contribution mass instead of a share of the tiles", True)
_reg("DGT_CONTRIBUTION_ALPHA", 0.3, float, "EMA factor of block contribution")
_reg("DGT_BLOCK_SIZE", 4096, int, "bytes per block")
_reg("DGT_INFO", 0, int, "log one line per DGT-split push (key, blocks, effective k, blocks sent)", True)
_reg("ENABLE_INTER_TS", 0, int, "TSEngine between parties")
_reg("ENABLE_INTRA_TS", 0, int, "TSEngine inside a party")
_reg("MAX_GREED_RATE_TS", 0.9, float, "ε-greedy cap")
# engine / storage / misc ------------------------------------------------------------------------------------
_reg("MXNET_ENGINE_TYPE", "ThreadedEnginePerDevice", str, "NaiveEngine serialises everything")
_reg("MXNET_CPU_WORKER_NTHREADS", 1, int, "")
_reg("MXNET_CPU_PRIORITY_NTHREADS", 4, int, "")
_reg("MXNET_GPU_WORKER_NTHREADS", 2, int, "")
_reg("MXNET_GPU_COPY_NTHREADS", 2, int, "")
_reg("MXNET_GPU_MEM_POOL_TYPE", "Naive", str, "Naive | Round | Unpooled: bucketing of the native device pool (storage.DevicePool, csrc/kernels/storage_gpu.cu)")
_reg("MXNET_GPU_MEM_POOL_RESERVE", 5, int, "percent of device memory the native pool keeps free (cached blocks are released first)")
_reg("MXNET_GPU_MEM_POOL_PAGE_SIZE", 4096, int, "bucket granularity of the native device pool")
_reg("MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF", 24, int, "log2 size above which the Round pool uses power-of-two buckets")
_reg("GEOMX_GPU_MEM_POOL", "torch", str, "torch | native: native installs the pool as PyTorch's CUDA allocator at import")
_reg("PS_BUFFER_POOL_MB", 1024, int, "cache of released message buffers >= 64 KiB (csrc/hips/block_pool.h); 0 disables pooling")
_reg("GEOMX_SERVER_THREADS", 0, int, "threads of a server's optimizer step / aggregation on big tensors; 0 = min(4, cores/2)")
_reg("MXNET_PROFILER_AUTOSTART", 0, int, "")
_reg("MXNET_PROFILER_MODE", 0, int, "")
_reg("MXNET_ENFORCE_DETERMINISM", 0, int, "")
# geomx_b200 additions ---------------------------------------------------------------------------------------
_reg("GEOMX_FABRIC", "auto", str, "auto | symm (NVSwitch symmetric memory) | nccl (oracle) | tcp")
_reg("GEOMX_EMULATE_DELAY_MS", 0, int, "emulated one-way latency (ms) of data messages on the global plane (WAN between parties)")
_reg("GEOMX_FUSED_TIER_PULL", 1, int, "global server answers a local server's dense push with the fresh value (one inter-party round trip per key instead of two)")
_reg("GEOMX_INLINE_RESPONSES", 1, int, "TCP plane: workers handle responses on the receive thread (one wake-up less per message); 0 = queue them")
_reg("GEOMX_SERVER_CKPT_PREFIX", "", str, "servers write <prefix>.server<rank>{g,l} every GEOMX_SERVER_CKPT_EVERY rounds")
_reg("GEOMX_SERVER_CKPT_EVERY", 0, int, "rounds between periodic server-state checkpoints (0 = off)")
_reg("GEOMX_SERVER_RESUME", 0, int, "a (re)started global / stand-alone server adopts the checkpoint under GEOMX_SERVER_CKPT_PREFIX")
_reg("GEOMX_NUM_PARTIES", 0, int, "fabric mode: number of parties (0 → DMLC_NUM_GLOBAL_WORKER, else 2 when the world is even, else 1)")
_reg("GEOMX_FP8_TRANSPORT", 0, int, "block-scaled e4m3 payload on the fp16/MPQ path")
_reg("GEOMX_SYNTHETIC_SIZE", 0, int, "shrink synthetic datasets")


def get(name):
    default, typ, _, _ = _F[name]
    v = os.environ.get(name)
    if v is None or v == "":
        return default
    try:
        return typ(float(v)) if typ is int else typ(v)
    except ValueError:
        return default


def describe():
    rows = ["%-42s %-14s %s" % ("name", "default", "doc")]
    for k, (d, t, doc, h) in _F.items():
        rows.append("%-42s %-14s %s%s" % (k, d, doc, "" if h else " (parsed, unused — as in the reference)"))
    return "\n".join(rows)


def names():
    return list(_F)
