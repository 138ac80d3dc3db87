"""ctypes loader for libgalah_b200.so (the C ABI declared in include/galah_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 device can be
bound, every compute call raises.
"""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libgalah_b200.so")


class GalahB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[galah_b200 rc={code}] {message}")
        self.code = code
        self.message = message


class Pair(ctypes.Structure):
    _fields_ = [("i", ctypes.c_uint32), ("j", ctypes.c_uint32), ("common", ctypes.c_uint32),
                ("total", ctypes.c_uint32), ("ani", ctypes.c_float)]


class Clusters(ctypes.Structure):
    _fields_ = [("members", ctypes.POINTER(ctypes.c_uint32)), ("offsets", ctypes.POINTER(ctypes.c_uint64)),
                ("n_clusters", ctypes.c_size_t), ("ani_calls", ctypes.c_uint64),
                ("n_preclusters", ctypes.c_uint32), ("largest_precluster", ctypes.c_uint32)]


class AniResult(ctypes.Structure):
    _fields_ = [("ani", ctypes.c_float), ("af_query", ctypes.c_float), ("af_ref", ctypes.c_float),
                ("estimator", ctypes.c_uint32), ("sum_fx", ctypes.c_uint64), ("n_chunks", ctypes.c_uint32),
                ("sum_m", ctypes.c_uint32), ("span_m", ctypes.c_uint32), ("span_n", ctypes.c_uint32),
                ("n_chains", ctypes.c_uint32), ("cov_q", ctypes.c_uint32), ("cov_r", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


class ClusterStats(ctypes.Structure):
    _fields_ = [("n_precluster_hits", ctypes.c_uint64), ("n_ani_pairs", ctypes.c_uint64),
                ("ani_chain_ms", ctypes.c_float), ("ingest_ms", ctypes.c_float), ("sketch_ms", ctypes.c_float),
                ("index_ms", ctypes.c_float), ("prefilter_ms", ctypes.c_float), ("ani_ms", ctypes.c_float),
                ("engine_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("ani_waves", ctypes.c_uint32)]


ANI_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                          ctypes.POINTER(ctypes.c_float))
ANI_BATCH_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float))

_lib = None

u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
f32p = ctypes.POINTER(ctypes.c_float)
pairpp = ctypes.POINTER(ctypes.POINTER(Pair))
sizep = ctypes.POINTER(ctypes.c_size_t)
strp = ctypes.POINTER(ctypes.c_char_p)
vp = ctypes.c_void_p

_SIGNATURES = {
    "galah_b200_init": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_device_count": (ctypes.c_int, []),
    "galah_b200_last_error": (ctypes.c_char_p, []),
    "galah_b200_version": (ctypes.c_char_p, []),
    "galah_b200_free": (None, [vp]),
    "galah_b200_launch_count": (ctypes.c_uint64, []),
    "galah_b200_prefilter_mode": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_prefilter_last_timing": (ctypes.c_int, [f32p, f32p]),
    "galah_b200_prefilter_stream_chunks": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_prefilter_last_host_timing": (ctypes.c_int, [f32p]),
    "galah_b200_sketch_files": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_uint8, ctypes.c_uint32,
                                               ctypes.c_uint64, ctypes.c_int, u64p, u32p]),
    "galah_b200_sketch_packed": (ctypes.c_int, [u32p, u32p, u64p, ctypes.c_size_t, ctypes.c_uint8,
                                                ctypes.c_uint32, ctypes.c_uint64, u64p, u32p]),
    "galah_b200_sketch_packed_device": (ctypes.c_int, [vp, vp, vp, ctypes.c_size_t, ctypes.c_uint8,
                                                       ctypes.c_uint32, ctypes.c_uint64, vp, vp, vp]),
    "galah_b200_prefilter": (ctypes.c_int, [u64p, u32p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                            ctypes.c_float, pairpp, sizep]),
    "galah_b200_prefilter_shard": (ctypes.c_int, [u64p, u32p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                                  ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, pairpp, sizep]),
    "galah_b200_prefilter_device": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                                   ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, vp,
                                                   pairpp, sizep]),
    "galah_b200_prefilter_enqueue": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                                    ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32,
                                                    ctypes.c_int, vp, vp, ctypes.c_size_t, vp]),
    "galah_b200_finish_candidates": (ctypes.c_int, [vp, ctypes.c_size_t, ctypes.c_uint8, ctypes.c_float, pairpp, sizep]),
    "galah_b200_blocklist_layout": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_size_t, sizep, sizep, sizep]),
    "galah_b200_blocklist_build": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                                  ctypes.c_size_t, vp, vp, vp, vp, vp]),
    "galah_b200_prefilter_join_items_enqueue": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                                               ctypes.c_float, vp, vp, vp, vp, vp, ctypes.c_size_t,
                                                               ctypes.c_int, vp, vp, ctypes.c_size_t, vp]),
    "galah_b200_table_max_device": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "galah_b200_blocklist_build_local": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, ctypes.c_size_t,
                                                        vp, vp, vp, vp, vp]),
    "galah_b200_prefilter_join_enqueue": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint8,
                                                         ctypes.c_float, vp, vp, vp, vp, ctypes.c_uint32,
                                                         ctypes.c_uint32, vp, vp, ctypes.c_size_t, vp]),
    "galah_b200_finch_distances": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_uint32,
                                                  ctypes.c_uint8, ctypes.c_int, pairpp, sizep]),
    "galah_b200_ani_index_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(vp)]),
    "galah_b200_ani_index_reserve": (ctypes.c_int, [vp, ctypes.c_size_t]),
    "galah_b200_ani_finish": (ctypes.c_int, [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_float, vp]),
    "galah_b200_chunk_identity_fx": (ctypes.c_uint64, [ctypes.c_uint32, ctypes.c_uint32]),
    "galah_b200_print2_parse_f32": (ctypes.c_float, [ctypes.c_double]),
    "galah_b200_ani_index_free": (None, [vp]),
    "galah_b200_ani_index_add_files": (ctypes.c_int, [vp, strp, ctypes.c_size_t, ctypes.c_int]),
    "galah_b200_ani_index_add_packed": (ctypes.c_int, [vp, u32p, u32p, u64p, ctypes.c_size_t, u64p, u32p, u32p]),
    "galah_b200_ani_index_add_packed_device": (ctypes.c_int, [vp, vp, vp, vp, u64p, u64p, ctypes.c_size_t, vp]),
    "galah_b200_ani_index_size": (ctypes.c_size_t, [vp]),
    "galah_b200_ani_index_clear": (ctypes.c_int, [vp]),
    "galah_b200_ani_index_genome": (ctypes.c_int, [vp, ctypes.c_size_t, u64p, u32p, u64p]),
    "galah_b200_ani_index_seeds": (ctypes.c_int, [vp, ctypes.c_size_t, u32p, u32p, u32p, ctypes.c_size_t]),
    "galah_b200_ani_pairs": (ctypes.c_int, [vp, u32p, ctypes.c_size_t, ctypes.c_float, ctypes.c_int,
                                            ctypes.POINTER(AniResult)]),
    "galah_b200_ani_last_timing": (ctypes.c_int, [vp, f32p, f32p]),
    "galah_b200_cluster_from_ani_table": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                         ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]),
    "galah_b200_cluster_from_ani_tables": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]),
    "galah_b200_cluster_from_distances": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_void_p,
                                                         ctypes.c_size_t, ctypes.c_int, ctypes.c_float,
                                                         ANI_FN, vp, ctypes.POINTER(Clusters)]),
    "galah_b200_cluster_from_distances_batched": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                                 ctypes.c_float, ANI_BATCH_FN, vp, ctypes.c_uint32,
                                                                 ctypes.POINTER(Clusters), ctypes.POINTER(ctypes.c_uint32)]),
    "galah_b200_cluster_lazy": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_clusters_free": (None, [ctypes.POINTER(Clusters)]),
    "galah_b200_cluster_files": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                ctypes.c_int, ctypes.c_int, ctypes.POINTER(Clusters),
                                                ctypes.POINTER(ClusterStats)]),
    "galah_b200_cluster_packed": (ctypes.c_int, [vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                                                 ctypes.c_float, ctypes.c_int, ctypes.POINTER(Clusters),
                                                 ctypes.POINTER(ClusterStats)]),
    "galah_b200_cluster_packed_device": (ctypes.c_int, [vp, vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_float,
                                                        ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                        ctypes.POINTER(Clusters), ctypes.POINTER(ClusterStats)]),
    "galah_b200_init_devices": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_cluster_packed_multi": (ctypes.c_int, [vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float,
                                                       ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                       ctypes.POINTER(Clusters), ctypes.POINTER(ClusterStats)]),
    "galah_b200_cluster_files_multi": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                      ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Clusters),
                                                      ctypes.POINTER(ClusterStats)]),
    "galah_b200_session_create": (ctypes.c_int, [ctypes.POINTER(vp)]),
    "galah_b200_session_free": (None, [vp]),
    "galah_b200_session_set_clusterer": (ctypes.c_int, [vp, ctypes.c_int]),
    "galah_b200_finch_method_name": (ctypes.c_char_p, []),
    "galah_b200_skani_method_name": (ctypes.c_char_p, []),
    "galah_b200_session_finch_distances": (ctypes.c_int, [vp, strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_uint32,
                                                          ctypes.c_uint8, ctypes.c_int, ctypes.c_int, pairpp, sizep]),
    "galah_b200_session_finch_distances_contigs": (ctypes.c_int, [vp, strp, ctypes.c_size_t, strp, ctypes.c_size_t, pairpp, sizep]),
    "galah_b200_session_finch_distances_with_references": (ctypes.c_int, [vp, strp, ctypes.c_size_t, strp, ctypes.c_size_t,
                                                                          pairpp, sizep]),
    "galah_b200_session_skani_distances": (ctypes.c_int, [vp, strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, pairpp, sizep]),
    "galah_b200_session_skani_distances_contigs": (ctypes.c_int, [vp, strp, ctypes.c_size_t, strp, ctypes.c_size_t,
                                                                  ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                                  ctypes.c_int, pairpp, sizep]),
    "galah_b200_session_skani_distances_with_references": (ctypes.c_int, [vp, strp, ctypes.c_size_t, strp, ctypes.c_size_t,
                                                                          ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                                          ctypes.c_int, pairpp, sizep]),
    "galah_b200_session_calculate_ani": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_float, ctypes.c_int,
                                                        f32p, ctypes.POINTER(ctypes.c_int)]),
    "galah_b200_session_stats": (ctypes.c_int, [vp, u64p, u64p, u64p]),
    "galah_b200_contig_names": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.POINTER(ctypes.POINTER(ctypes.c_char_p)), sizep]),
    "galah_b200_contig_names_free": (None, [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t]),
    "galah_b200_ingest_packed": (ctypes.c_int, [vp, vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_int, vp, vp, vp, f32p]),
    "galah_b200_ingest_packed_sparse": (ctypes.c_int, [vp, u64p, u64p, ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, vp, vp, vp, f32p]),
    "galah_b200_cluster_packed_sparse": (ctypes.c_int, [vp, u64p, u64p, ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_float,
                                                        ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.POINTER(Clusters),
                                                        ctypes.POINTER(ClusterStats)]),
    "galah_b200_ingest_packed_markers": (ctypes.c_int, [vp, vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32,
                                                        vp, vp, vp, f32p]),
    "galah_b200_marker_row_capacity": (ctypes.c_uint32, [ctypes.c_uint64, ctypes.c_int]),
    "galah_b200_prefilter_join_enqueue_screen": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, vp, vp,
                                                                vp, vp, ctypes.c_uint32, ctypes.c_uint32, vp, vp,
                                                                ctypes.c_size_t, vp]),
    "galah_b200_ani_index_export_tables": (ctypes.c_int, [vp, vp, u64p, u64p]),
    "galah_b200_ani_index_attach_peer": (ctypes.c_int, [vp, vp, u64p, u64p, ctypes.c_size_t, u32p]),
    "galah_b200_skani_distances": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, pairpp, sizep, sizep]),
    "galah_b200_skani_distances_multi": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, pairpp, sizep, sizep]),
    "galah_b200_cluster_files_skani_multi": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                            ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                            ctypes.POINTER(Clusters), ctypes.POINTER(ClusterStats)]),
    "galah_b200_skani_distances_packed_device": (ctypes.c_int, [vp, vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_float,
                                                                ctypes.c_float, ctypes.c_int, ctypes.c_int, vp,
                                                                ctypes.POINTER(ctypes.POINTER(Pair)), sizep,
                                                                ctypes.POINTER(ctypes.c_uint64), f32p]),
    "galah_b200_skani_distances_packed_multi": (ctypes.c_int, [vp, vp, u64p, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float,
                                                               ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                               ctypes.POINTER(ctypes.POINTER(Pair)), sizep,
                                                               ctypes.POINTER(ctypes.c_uint64)]),
    "galah_b200_cluster_files_skani": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_float, ctypes.c_float,
                                                      ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.POINTER(Clusters), ctypes.POINTER(ClusterStats)]),
    "galah_b200_device_ingest": (ctypes.c_int, [ctypes.c_int]),
    "galah_b200_decode_fasta_device": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t),
                                                      ctypes.c_size_t] + [ctypes.POINTER(u32p)] * 2 +
                                       [ctypes.POINTER(u64p)] * 7 + [f32p]),
    "galah_b200_pack_fasta_file": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(u32p), ctypes.POINTER(u32p), u64p,
                                                  ctypes.POINTER(u64p), ctypes.POINTER(u64p), sizep]),
    "galah_b200_genome_stats": (ctypes.c_int, [strp, ctypes.c_size_t, ctypes.c_int, vp]),
    "galah_b200_synth_packed_device": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_size_t,
                                                      ctypes.c_uint64, vp, vp, vp, vp]),
    "galah_b200_synth_packed_device_ex": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_size_t,
                                                         ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp]),
    "galah_b200_stream": (ctypes.c_void_p, []),
}


def exported_symbols():
    """Every symbol include/galah_b200.h declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GalahB200Error(
                -1, f"{LIB_PATH} is missing: build it with `python -m galah_b200.build` "
                    "(there is no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise GalahB200Error(rc, lib().galah_b200_last_error().decode("utf-8", "replace"))
