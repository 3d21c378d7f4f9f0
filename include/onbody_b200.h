/*
 * onbody_b200.h - C ABI of the B200-native (sm_100a) summation hot path of onbody.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types. The shared library
 * (onbody_b200/libonbody_b200.so) links the CUDA runtime statically, so callers do not link CUDA.
 * There is NO CPU fallback: every entry point fails with a non-zero code (and onb_error() says why)
 * when no sm_100 device is usable.
 *
 * The entry points mirror, one for one, the free functions the reference's drivers call between
 * "particles are initialised" and "results are compared" (reference src/ongrav3d.cpp:600-908):
 *
 *   onb_make_tree     <- makeTree()                   barneshut.hpp:814-854  (splitNode :594, finishTree :717)
 *   onb_refine        <- refineTree()                 barneshut.hpp:901-936
 *   onb_upward        <- calcBarycentricLagrange()    BarycentricLagrange.hpp:255-417 (+ eq*.resize, ongrav3d.cpp:645,696)
 *   onb_naive         <- nbody_naive()                barneshut.hpp:46-53
 *   onb_treecode1     <- nbody_treecode1()            barneshut.hpp:107-132
 *   onb_treecode2     <- nbody_treecode2()            barneshut.hpp:189-222  (pointwise)
 *   onb_treecode3     <- nbody_treecode3()            barneshut.hpp:299-337  (boxwise)
 *   onb_fastsumm      <- nbody_fastsumm()             ongrav3d.cpp:206-452   (dual-tree, incl. calcBarycentricDownward)
 *   onb_zero_vels     <- Parts::zero_vels()           Parts.hpp:179-183
 *
 * The drop-in Fortran-style entry points external_vel_solver_f_ / external_vel_direct_f_ live in two
 * separate shim libraries (same symbol names, different arity - exactly as in the reference):
 * include/onbody_bh2dvort.h and include/onbody_bh3dvortgrads.h.
 *
 * Beyond the reference's single-device, float-accumulating drivers the same ABI carries: the multi-GPU communicator
 * (onb_comm_*: with it attached the phase calls above run distributed and return the same bits), the lean memory mode for
 * N ~ 1e9 (onb_set_memory_mode), and the reference's ACCUM = double build option (onb_set_accum, ongrav3d.cpp:7-8).
 *
 * "which" arguments: 0 = sources, 1 = targets, 2 = equivalent sources, 3 = equivalent targets.
 * All array arguments are HOST pointers; planar layout: x is [PD][n], s is [SD][n], u is [OD][n].
 */
#ifndef ONBODY_B200_H
#define ONBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONB_API __attribute__((visibility("default")))

typedef struct onb_context onb_context;

/* physics of the pair kernel = which reference driver is being replaced */
enum {
    ONB_GRAV3D     = 0,  /* ongrav3d.cpp:44-58       PD 3 SD 1 OD 3   19 flop/pair */
    ONB_VORT3D     = 1,  /* onvort3d.cpp:44-59       PD 3 SD 3 OD 3   28 flop/pair */
    ONB_VORTGRAD3D = 2,  /* onvortgrad3d.cpp:45-76   PD 3 SD 3 OD 12  64 flop/pair */
    ONB_VORT2D     = 3,  /* interface2dvort.cpp:39-50  PD 2 SD 1 OD 2 13 flop/pair */
    ONB_VORT2DTR   = 4   /* onvort2d.cpp:44-55 (+target radius) PD 2 SD 1 OD 2 15 flop/pair */
};

/* arithmetic of the pair / interpolation kernels */
enum {
    ONB_ARITH_FAST   = 0, /* rsqrt + FMA: the product path */
    ONB_ARITH_STRICT = 1  /* the reference's IEEE operation sequence, for bit-exact verification */
};

/* error codes */
enum {
    ONB_OK = 0,
    ONB_ERR_CUDA = 1,       /* a CUDA call failed (no device, OOM, launch failure) */
    ONB_ERR_ARG = 2,        /* invalid argument or call order */
    ONB_ERR_CAPACITY = 3,   /* an internal work queue overflowed */
    ONB_ERR_UNSUPPORTED = 4 /* a path the GPU build does not implement (e.g. order < 1) */
};

/* lifetime ----------------------------------------------------------------------------------- */
/* creates a context on CUDA device `device`; returns NULL (see onb_last_create_error) if there is none */
ONB_API onb_context* onb_create(int physics, int device);
ONB_API void         onb_destroy(onb_context* c);
ONB_API const char*  onb_error(const onb_context* c);
ONB_API const char*  onb_last_create_error(void);

/* block size (-b, default 128), barycentric order (-o, >=1) and arithmetic; call before set_sources */
ONB_API int onb_set_params(onb_context* c, int block_size, int order, int arith);
/* the per-pair flop constant used in the returned flop estimates (defaults: the drivers' nbody_kernel_flops();
 * interface2dvorttr.cpp:53 counts 13 where onvort2d.cpp:57 counts 15 for the same kernel) */
ONB_API void onb_set_flops_per_pair(onb_context* c, int flops);
ONB_API void onb_dims(const onb_context* c, int* pd, int* sd, int* od, int* has_fastsumm);

/* inputs (copies into the context's own device arrays; the pointers may be host memory - pinned for full PCIe
 * speed - or device memory, the copy kind is resolved by unified addressing) ------------------------------ */
ONB_API int onb_set_sources(onb_context* c, uint64_t n, const float* x, const float* r, const float* s);
ONB_API int onb_set_targets(onb_context* c, uint64_t n, const float* x, const float* r);
/* the same with one pointer per plane (x[PD], s[SD]): the Fortran-style entry points hand over separate arrays
 * (interface3dvortgrads.cpp:247-254), which then need no re-packing on the host */
ONB_API int onb_set_sources_planes(onb_context* c, uint64_t n, const float* const* x, const float* r, const float* const* s);
ONB_API int onb_set_targets_planes(onb_context* c, uint64_t n, const float* const* x, const float* r);
/* opt-in: with on != 0, set_sources / set_targets return as soon as the copies from PINNED host (or device) buffers
 * are enqueued; the source copy runs on the context's stream and the target copy on its second stream, so that
 * onb_make_trees builds the source tree while the targets are still crossing PCIe. The caller must leave the
 * buffers untouched until the next phase call that consumes them has returned (every phase call blocks). Pageable
 * buffers are still copied synchronously. Default off: the copies complete before set_* returns. */
ONB_API int onb_set_async_inputs(onb_context* c, int on);
/* opt-in, with a communicator attached (every context is handed the SAME arrays): each context copies only its 1/nranks
 * slice of every plane from the caller's buffers and the slices are replicated by one grouped all-gather over NVLink, so
 * that every input byte crosses PCIe once per node instead of once per GPU. The device arrays end up identical. */
ONB_API int onb_set_sliced_inputs(onb_context* c, int on);
/* the drivers' own synthetic initialisation (std::mt19937(12345), Parts.hpp:99-109,169-176), done on the
 * host into caller buffers: x [PD][n], r [n], s [SD][n] (s may be NULL for targets). strength_mode 1 = wave_strengths */
ONB_API int onb_driver_inputs(int physics, uint64_t n, int strength_mode, float* x, float* r, float* s);

/* phases ------------------------------------------------------------------------------------- */
ONB_API int onb_make_tree(onb_context* c, int which);
/* multi-GPU: build only the part of the tree that overlaps the tree-order particle range [lo,hi) (leaf aligned, see
 * onb_shard_particle_range). Ancestors of the range are split in full, siblings outside it are left unsorted; the node
 * index/count arrays are complete (they depend on n and the block size only). For TARGETS that is all a rank needs.
 * For SOURCES the ranks then exchange their ranges of the particle planes (onb_device_ptr + an NCCL all-gather issued by
 * the host language) and call onb_finish_tree, which recomputes every node array bottom-up, bit-identical to a full build. */
ONB_API int onb_make_tree_range(onb_context* c, int which, uint64_t lo, uint64_t hi);
ONB_API int onb_finish_tree(onb_context* c, int which);
/* source and target tree in one call: the two (independent) builds are enqueued on two streams and overlap on the device;
 * results are identical to two onb_make_tree[_range] calls */
ONB_API int onb_make_trees(onb_context* c);
ONB_API int onb_make_trees_range(onb_context* c, uint64_t src_lo, uint64_t src_hi, uint64_t tgt_lo, uint64_t tgt_hi);
/* everything between the tree builds and onb_fastsumm in one call, source side and target side overlapped on two streams:
 * [onb_finish_tree(0)] onb_upward(0)  |  [onb_finish_tree(1)] onb_refine(1) onb_upward(1)
 * (ongrav3d.cpp:636-724 runs the two independent chains one after the other). finish != 0: the multi-GPU variant after the
 * plane exchange, with the in-leaf refinement restricted to the target range [tgt_lo, tgt_hi). Results are identical to
 * the separate calls. */
ONB_API int onb_prepare_eval(onb_context* c, int finish, uint64_t tgt_lo, uint64_t tgt_hi);
/* restrict the following onb_refine to the leaves of [lo,hi) again (onb_finish_tree resets the range to the whole set) */
ONB_API int onb_set_build_range(onb_context* c, int which, uint64_t lo, uint64_t hi);
/* the tree-order particle range of shard `rank` of `nranks` for a set of n particles (contiguous leaves) */
/* the same without a context (pure host arithmetic) */
ONB_API int onb_shard_range_for(uint64_t n, int block, int rank, int nranks, uint64_t* lo, uint64_t* hi);
ONB_API int onb_shard_particle_range(const onb_context* c, uint64_t n, int rank, int nranks, uint64_t* lo, uint64_t* hi);
ONB_API int onb_refine(onb_context* c, int which);
ONB_API int onb_upward(onb_context* c, int which);
ONB_API int onb_zero_vels(onb_context* c);
ONB_API int onb_naive(onb_context* c, uint64_t tskip, float* flops);
ONB_API int onb_treecode1(onb_context* c, float theta, float* flops);
ONB_API int onb_treecode2(onb_context* c, float theta, float* flops);
ONB_API int onb_treecode3(onb_context* c, float theta, float* flops);
ONB_API int onb_fastsumm(onb_context* c, float theta);

/* target sharding for multi-GPU runs: this context evaluates only the targets of shard `rank` of `nranks`
 * (contiguous tree-order ranges of target leaves, ceil(leaves/nranks) each); without a communicator the trees themselves
 * are built in full (or by the explicit *_range calls above). */
ONB_API int onb_set_shard(onb_context* c, int rank, int nranks);

/* multi-GPU communicator ---------------------------------------------------------------------------------------
 * (the reference has no multi-device path; north_star: targets sharded, source tree + equivalent particles replicated by
 * NCCL all-gathers over NVLink.) One context per GPU. With a communicator attached the SAME phase calls run distributed:
 * onb_make_tree(s) sorts only this rank's leaf range and completes the node arrays from exchanged per-leaf records,
 * onb_upward(0) / onb_prepare_eval anterpolate the rank's own nodes and exchange the strengths, onb_refine / onb_upward(1) /
 * every evaluation method work on the rank's target shard; results are bit-identical to the single-GPU run. Every context
 * must be given the same inputs and must make the same calls in the same order. NCCL is loaded with dlopen at first use
 * (ONB_NCCL_LIB overrides "libnccl.so.2").
 *   one process per GPU : rank 0 calls onb_comm_unique_id, ships the 128 bytes to the others, all call onb_comm_init_rank
 *   one process, n GPUs : onb_comm_init_all(ctxs, n), then one host thread per context makes the phase calls
 *   onb_comm_init_loopback: contexts of one process joined by device copies instead of NCCL - on ANY devices, also all on
 *   the same one; exists so that the distributed path can be verified on a single GPU, one host thread per context. */
ONB_API int onb_comm_unique_id(void* id, uint64_t bytes);
ONB_API int onb_comm_init_rank(onb_context* c, int rank, int nranks, const void* id, uint64_t bytes);
ONB_API int onb_comm_init_all(onb_context** ctxs, int n);
ONB_API int onb_comm_init_loopback(onb_context** ctxs, int n);
ONB_API int onb_comm_destroy(onb_context* c);
/* transport: 0 none, 1 NCCL, 2 loopback */
ONB_API int onb_comm_info(const onb_context* c, int* rank, int* nranks, int* transport, int* nccl_version);
/* the partition arithmetic, pure host code (no GPU needed): particles per rank, and per tree level the node-id intervals
 * [own_lo, own_hi) inside / [need_lo, need_hi) overlapping the rank's range plus the nodes straddling rank boundaries
 * (shared[level * nranks + k], k < nshared[level]). Returns the number of levels, or a negative error code. */
ONB_API uint64_t onb_shard_chunk_for(uint64_t n, int block, int nranks);
ONB_API int onb_plan_query(uint64_t n, int block, int nranks, int rank, int max_levels, uint32_t* own_lo, uint32_t* own_hi,
                           uint32_t* need_lo, uint32_t* need_hi, uint32_t* nshared, uint32_t* shared);

/* precision of the accumulations - the reference's compile-time ACCUM macro (ongrav3d.cpp:7-8; README.md:107-112: "performing
 * accumulations in fp64 allows the RMS error to drop to about 4e-7"). 0 (default) = float like the shipped drivers; 1 = double:
 * the pair arithmetic stays fp32 (STORE = float), every "+=" into a target value and the downward interpolation run in fp64 and
 * the outputs are kept in fp64 (onb_get_results_f64; onb_get_parts / onb_add_results_* return them rounded to float). All
 * methods (direct, treecode1/2/3, dual tree), all physics; bit-exact against the reference built with ACCUM = double in
 * ONB_ARITH_STRICT. Call before set_targets. */
ONB_API int onb_set_accum(onb_context* c, int accum_double);
/* which = 1 targets, 3 equivalent target points; u is [OD][n] doubles */
ONB_API int onb_get_results_f64(onb_context* c, int which, double* u);

/* memory mode: ONB_MEM_LEAN trades allocator calls for footprint (N = 1e9 on 8 GPUs, BASELINE configs[4]): one tree build at
 * a time and its scratch returned to the driver, the SoA source planes released once the float4 tiles exist (onb_get_parts
 * of sources then fails), and - in a sharded run - the target outputs and the equivalent target points become sparse planes
 * backed by memory only where this rank's shard touches them. Call before set_sources / set_targets. */
enum { ONB_MEM_NORMAL = 0, ONB_MEM_LEAN = 1 };
ONB_API int onb_set_memory_mode(onb_context* c, int mode);
/* bytes currently allocated on the context's device by this process and the device total (cudaMemGetInfo) */
ONB_API int onb_device_memory(onb_context* c, uint64_t* used, uint64_t* total);

/* outputs (device -> host copies) ------------------------------------------------------------ */
ONB_API uint64_t onb_count(const onb_context* c, int which);
ONB_API int onb_get_parts(onb_context* c, int which, float* x, float* r, float* s, float* u, uint64_t* gidx);
/* results scattered back to the caller's original target order and ADDED to u (the reference's
 * external_vel_solver_f_ semantics, interface3dvortgrads.cpp:384-395); u is [OD][n] */
ONB_API int onb_add_results_original_order(onb_context* c, float* u);
/* the same with one pointer per output plane (out[OD]); host or device pointers. Device pointers are updated in place by one
 * kernel; host pointers through one scatter kernel and a pinned double buffer whose copies overlap the host's += */
ONB_API int onb_add_results_planes(onb_context* c, float* const* out);
/* multi-GPU: the output planes of this context's shard only - u[d*plane_stride + i] for i in [*lo,*hi), tree order
 * (plane_stride 0 = n: the full [OD][n] layout, every rank filling its own part of one shared array) */
ONB_API int onb_get_shard_results(onb_context* c, float* u, uint64_t plane_stride, uint64_t* lo, uint64_t* hi);
ONB_API int onb_tree_shape(const onb_context* c, int which, int* levels, int* numnodes);
ONB_API int onb_get_tree(onb_context* c, int which, float* x, float* nc, float* ns, float* nr, float* pr, float* s,
                 uint64_t* ioffset, uint64_t* num, uint64_t* epoffset, uint64_t* epnum);

/* counters of the last treecode / dual-tree call: sltp sbtp | sltl sbtl sltb sbtb tlc lpc bpc
 * (the reference's treecode_stats barneshut.hpp:58-60 and fastsumm_stats ongrav3d.cpp:193-196) */
ONB_API int onb_get_stats(const onb_context* c, uint64_t out[9]);

/* measurement ------------------------------------------------------------------------------- */
/* device time (CUDA events on the context's stream) of the named phase of the last call that ran it, in ms;
 * names: "tree", "refine", "upward", "lists", "p2p", "downward", "eval". Returns <0 if never run. */
ONB_API double   onb_phase_ms(const onb_context* c, const char* name);
/* step timer: CUDA events on the context's stream around everything issued between the two calls */
ONB_API int      onb_timer_start(onb_context* c);
ONB_API double   onb_timer_stop_ms(onb_context* c);
/* exact number of source-target pairs the pair kernels evaluated in the last evaluation call */
ONB_API uint64_t onb_last_pairs(const onb_context* c);
/* number of kernel launches issued by this context since creation */
ONB_API uint64_t onb_launch_count(const onb_context* c);
/* FP32 FMA issue-rate microbenchmark on this device: returns measured TFLOP/s (2 flop per FMA lane) */
ONB_API double   onb_measure_fp32_peak(onb_context* c);
/* raw device pointers for zero-copy plumbing (e.g. an NCCL all-gather issued from the host language):
 * field: 0..2 x[d], 3 r, 4..6 s[d], 7.. u[d]; returns NULL if absent */
ONB_API void*    onb_device_ptr(onb_context* c, int which, int field);

/* diagnostics of the last onb_make_tree / onb_refine: selects, partition passes, stall exits, elements scanned, and the
 * number of in-leaf sorts that met equal keys (where libstdc++'s introsort order had to be reproduced) */
ONB_API int onb_get_build_stats(onb_context* c, uint64_t out[5]);
/* pivot arithmetic of the partial select: 0 (default) = the reference source evaluated in IEEE order (matches the
 * reference built with -O2 -ffp-contract=off); 1 = with the contractions g++ -O3 -ffast-math applies to
 * barneshut.hpp:538-540 (matches that build's intra-leaf source order as well). Process-wide. */
ONB_API void onb_set_pivot_mode(int mode);

/* tuning / tests: targets per thread of the list-driven pair kernel (1, 2 or 4; 0 = built-in per-physics default). Process-wide. */
ONB_API void onb_set_p2p_tpt(int tpt);

/* test support: install an already-built tree + already-ordered particles (lets each phase be parity-
 * tested in isolation against the oracle). Arrays as in onb_get_tree. */
ONB_API int onb_load_tree(onb_context* c, int which, int levels, const float* x, const float* nc, const float* ns,
                  const float* nr, const float* pr, const float* s, const uint64_t* ioffset, const uint64_t* num);

#ifdef __cplusplus
}
#endif
#endif /* ONBODY_B200_H */
