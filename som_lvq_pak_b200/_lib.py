"""ctypes binding of the C ABI in include/bmu.h (som_lvq_pak_b200/libbmu_b200.so).

The library is the product; this module only declares its prototypes.  There is no CPU
fallback: if the shared library is missing the import of any compute entry point raises,
and on a machine without an sm_100 GPU every compute call returns an error code that
`check()` turns into a RuntimeError."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbmu_b200.so")

c_f = C.POINTER(C.c_float)
c_i32 = C.POINTER(C.c_int32)
c_i16 = C.POINTER(C.c_int16)
c_u8 = C.POINTER(C.c_ubyte)
c_d = C.POINTER(C.c_double)
c_i64 = C.POINTER(C.c_longlong)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/bmu.h declares
PROTOTYPES = {
    "bmu_init": (C.c_int, [C.c_int]),
    "bmu_shutdown": (None, []),
    "bmu_last_error": (C.c_char_p, []),
    "bmu_device_info": (C.c_int, [c_i32, c_i32, c_i32, C.POINTER(C.c_size_t)]),
    "bmu_set_search_path": (C.c_int, [C.c_int]),
    "bmu_launch_count": (C.c_long, []),
    "bmu_last_search_breakdown": (C.c_int, [C.POINTER(C.c_long)]),
    "bmu_last_search_kernel_ms": (C.c_int, [c_f]),
    "bmu_search_kernel_ms_history": (C.c_int, [C.c_int, c_f]),
    "bmu_codebook_create": (vp, [vp, C.c_long, C.c_int]),
    "bmu_codebook_create_dev": (vp, [vp, C.c_long, C.c_int]),
    "bmu_codebook_update": (C.c_int, [vp, vp]),
    "bmu_codebook_destroy": (None, [vp]),
    "bmu_search": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, vp, vp, vp]),
    "bmu_search_dev": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, vp, vp, vp, vp]),
    "bmu_search_stats_dev": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, C.c_long, vp, vp, vp, vp, vp,
                                       C.c_int, vp, vp]),
    "bmu_multi_init": (C.c_int, [C.c_int]),
    "bmu_multi_shards": (C.c_int, []),
    "bmu_multi_devices": (C.c_int, []),
    "bmu_multi_shard_bounds": (None, [C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "bmu_mcodebook_create": (vp, [vp, C.c_long, C.c_int]),
    "bmu_mcodebook_update": (C.c_int, [vp, vp]),
    "bmu_mcodebook_set_labels": (C.c_int, [vp, vp]),
    "bmu_mcodebook_destroy": (None, [vp]),
    "bmu_multi_search": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, vp, vp, vp, vp]),
    "bmu_comm_unique_id": (C.c_int, [vp]),
    "bmu_comm_init_rank": (C.c_int, [C.c_int, C.c_int, vp]),
    "bmu_comm_destroy": (C.c_int, []),
    "bmu_comm_broadcast_dev": (C.c_int, [vp, C.c_size_t, C.c_int, vp]),
    "bmu_comm_allreduce_stats_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp]),
    "bmu_host_alloc": (vp, [C.c_size_t]),
    "bmu_host_free": (None, [vp]),
    "bmu_host_register": (C.c_int, [vp, C.c_size_t]),
    "bmu_host_unregister": (C.c_int, [vp]),
    "bmu_set_copy_threads": (C.c_int, [C.c_int]),
    "bmu_som_train": (C.c_int, [vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                C.c_long, vp, vp, vp, vp, C.c_long]),
    "bmu_lvq_train": (C.c_int, [C.c_int, vp, vp, C.c_long, C.c_int, vp, vp, vp, C.c_long, vp, vp,
                                C.c_long, C.c_float, C.c_float, C.c_float, vp]),
    "bmu_trainer_create": (vp, [vp, C.c_long, C.c_int, vp, vp, C.c_long]),
    "bmu_trainer_set_som": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "bmu_trainer_set_lvq": (C.c_int, [vp, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, vp]),
    "bmu_trainer_steps": (C.c_int, [vp, vp, vp, vp, C.c_long]),
    "bmu_trainer_get_codes": (C.c_int, [vp, vp]),
    "bmu_trainer_get_unit_alpha": (C.c_int, [vp, vp]),
    "bmu_trainer_last_ms": (C.c_float, [vp]),
    "bmu_trainer_destroy": (None, [vp]),
    "bmu_qerror2": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, C.c_long, vp]),
    "bmu_class_nearest": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, vp, vp]),
    "bmu_identical_pairs": (C.c_int, [vp, vp, C.c_long, C.c_int, vp, C.c_long, vp]),
    "bmu_sammon": (C.c_int, [vp, vp, C.c_long, C.c_int, C.c_long, vp, vp, vp]),
    "bmu_rand_order": (None, [C.c_long, C.c_int, vp]),
    "bmu_sample_sequence": (None, [C.c_long, C.c_long, C.c_int, C.c_long, vp]),
    "bmu_som_schedule": (None, [C.c_long, C.c_long, C.c_long, C.c_float, C.c_float, C.c_int,
                                C.c_long, vp, vp, vp, vp, vp]),
    "bmu_replay_qerror": (C.c_float, [vp, vp, C.c_long, C.c_int]),
    "bmu_randinit_codes": (None, [vp, vp, C.c_long, C.c_int, C.c_long, C.c_int, vp]),
    "bmu_lvq_schedule": (None, [C.c_long, C.c_long, C.c_long, C.c_float, C.c_int, C.c_long, vp,
                                vp, vp]),
}



class Stats(C.Structure):
    """struct bmu_stats of include/bmu.h"""
    _fields_ = [("sum_sqrt", C.c_double), ("n_found", C.c_longlong), ("hist", vp), ("confusion", vp),
                ("sample_label", vp), ("n_labels", C.c_int)]


_lib = None


def _preload_env_nccl():
    """libbmu_b200 resolves NCCL at run time by its soname (multi-GPU only).  In a Python environment that
    ships its own libnccl.so.2 (the nvidia-nccl wheel PyTorch links against) that copy must be the one the
    process loads FIRST: two builds with the same soname cannot coexist, and a later `import torch` would be
    bound to whichever came first.  No-op when the wheel is absent (the system library is used then)."""
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            for f in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load():
    """dlopen the CUDA library (built in-tree by `make -C som_lvq_pak_b200/csrc`)."""
    global _lib
    if _lib is None:
        _preload_env_nccl()
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `make -C som_lvq_pak_b200/csrc` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("bmu error %d: %s" % (rc, load().bmu_last_error().decode()))
