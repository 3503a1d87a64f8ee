/*
 * parops.h -- C ABI of libparops.so, the B200 (sm_100a) spectral-operator apply engine
 * that replaces the array-type-dispatched apply path of ParametricOperators.jl.
 *
 * The reference has NO FFI boundary of its own: the interface is
 * Julia multiple dispatch on callable structs,
 *     (A::Op)(x::AbstractMatrix{D})            e.g. src/ParDFT.jl:26-32, src/ParKron.jl:190-220
 *     *(A, x)                                  src/ParOperator.jl:221-223
 *     adjoint(A), A(theta), init(A)            src/ParAdjoint.jl:11, src/ParOperator.jl:194-206
 * with the device picked by the array type (CuArray vs Array, src/ParCommon.jl:98-104).
 * The entry points below are what new Julia methods specialised on CuArray bind with
 * `ccall` (see INTEGRATION.md and parametricoperators.jl_b200/julia/ParametricOperatorsB200.jl):
 * the host lowers a (parameterised) operator tree into the flat `po_node` array, the
 * library compiles it into a plan of fused sm_100a kernels and applies it to raw device
 * pointers on the caller's stream.
 *
 * Conventions
 *  - every function returns a po_status (0 = ok, <0 = error); the message is available
 *    from po_last_error() (thread local).  Nothing throws or aborts across the ABI.  The
 *    Julia shim turns a non-zero status into `ParException` (src/ParCommon.jl:4-6).
 *  - arrays are column-major; `x` is (Domain, batch), `y` is (Range, batch); complex
 *    numbers are interleaved (re, im) pairs -- Julia's ComplexF32/ComplexF64 layout.
 *  - caller owns x, y, parameters, gradients and workspace; the library owns plans (DFT
 *    matrices, twiddles, index maps, box tables) until po_plan_destroy.  Inputs are const.
 *  - all device work is enqueued on the caller-supplied stream (cudaStream_t passed as
 *    void*); the library never calls cudaDeviceSynchronize on the apply path.
 */
#ifndef PAROPS_H
#define PAROPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PO_VERSION 100 /* 0.1.0, matches Project.toml:4 of the reference */

/* element types: the four the reference's tests exercise (test/types_and_sizes.jl:6-13) */
typedef enum po_dtype { PO_F32 = 0, PO_F64 = 1, PO_C64 = 2, PO_C128 = 3 } po_dtype;

typedef enum po_status {
  PO_OK = 0,
  PO_ERR_INVALID = -1,     /* bad argument / malformed tree                                  */
  PO_ERR_SHAPE = -2,       /* Domain/Range mismatch (the @assert of src/ParCompose.jl:16-19) */
  PO_ERR_DTYPE = -3,       /* DDT/RDT mismatch or unsupported element type                   */
  PO_ERR_CUDA = -4,        /* CUDA runtime error                                             */
  PO_ERR_NCCL = -5,        /* NCCL error / NCCL not loadable                                 */
  PO_ERR_UNSUPPORTED = -6, /* defined in the header, not applicable to this operator         */
  PO_ERR_WORKSPACE = -7    /* workspace too small                                            */
} po_status;

/* node kinds of the flat operator tree (one per ★ operator of SURVEY.md section 2) */
typedef enum po_node_kind {
  PO_NODE_IDENTITY = 0,    /* ParIdentity / ParDistributed(ParIdentity)   src/ParIdentity.jl:6-22      */
  PO_NODE_DFT = 1,         /* ParDFT; adjoint flag = inverse               src/ParDFT.jl:26-32          */
  PO_NODE_RESTRICTION = 2, /* ParRestriction; adjoint = zero-fill scatter  src/ParRestriction.jl:13-39  */
  PO_NODE_MATRIX = 3,      /* ParMatrix (parameterised); adjoint = theta'  src/ParMatrix.jl:42-46       */
  PO_NODE_DIAGONAL = 4,    /* ParDiagonal; adjoint = conj(theta)           src/ParDiagonal.jl:24-31     */
  PO_NODE_TENSOR = 5,      /* ParTensor ranks 4/5/6                        src/ParTensor.jl:35-113      */
  PO_NODE_KRON = 6,        /* ParKron with explicit application order      src/ParKron.jl:190-220       */
  PO_NODE_COMPOSE = 7,     /* ParCompose (children[0] applied last)        src/ParCompose.jl:44-56      */
  PO_NODE_REPARTITION = 8, /* ParRepartition                               src/ParRepartition.jl:140-192*/
  PO_NODE_BIAS = 9,        /* x .+ theta (m x 1 ParMatrix)                 src/ParMatrix.jl:53-55       */
  PO_NODE_GELU = 10        /* gelu.(x) of examples/fno.jl:43 (NNlib tanh form)                          */
} po_node_kind;

/*
 * One node of the flat tree.  `domain`/`range`/`ddt`/`rdt` describe the node AS APPLIED
 * (after the adjoint flag).  Per-kind payload in aux[aux_begin .. aux_begin+aux_count):
 *   DFT          [n]                        transform length (Domain of the un-adjointed op)
 *   RESTRICTION  [n, nranges, s1, e1, ...]  1-based inclusive ranges (src/ParRestriction.jl:5)
 *   MATRIX       [m, n]                     parameter is m x n column-major
 *   DIAGONAL     [n]
 *   TENSOR       [rank, weight_shape...]    rank 4: (ic,oc,nt,nxy); rank 5/6: (oc,ic,n...)
 *   KRON         order[N]                   1-based application order (src/ParKron.jl:33-57)
 *   REPARTITION  [N, global_size[N], dims_in[N], dims_out[N], rank]
 *   BIAS         [m]
 * children (KRON / COMPOSE) are child_index[child_begin .. child_begin+child_count).
 */
typedef struct po_node {
  int32_t kind;       /* po_node_kind                         */
  int32_t ddt, rdt;   /* po_dtype of the domain / range       */
  int32_t adjoint;    /* 1: this leaf is wrapped in ParAdjoint */
  int32_t param_slot; /* index into params[], -1 if none      */
  int32_t child_begin, child_count;
  int32_t aux_begin, aux_count;
  int64_t domain, range;
} po_node;

typedef struct po_ctx po_ctx;
typedef struct po_plan po_plan;

/* plan flags */
#define PO_PLAN_DEFAULT 0u
#define PO_PLAN_NO_FUSION 1u    /* one kernel per reference leaf apply, reference order (debug / A-B runs) */
#define PO_PLAN_NO_FASTPATH 2u  /* generic mode-product kernels only (no shape-specialised fused kernels)  */
#define PO_PLAN_NO_PEER 8u      /* ParRepartition through pack + grouped ncclSend/ncclRecv + unpack instead of the one-kernel NVLink peer-store exchange */
#define PO_PLAN_NO_PIPELINE 4u  /* fused kernels: one unit per CTA instead of the persistent warp-specialised form */

/* ---- library / context -------------------------------------------------------------- */
int32_t po_version(void);
const char* po_last_error(void);
const char* po_status_string(int32_t status);
int32_t po_ctx_create(int32_t device, po_ctx** out);
int32_t po_ctx_destroy(po_ctx* ctx);
int32_t po_ctx_device(const po_ctx* ctx, int32_t* device);
/* Device-side waits (bulk-copy barriers of the streaming kernels, peer flags of the NVLink ParRepartition) are bounded by
 * time (PO_PEER_TIMEOUT_MS, default 600000; <= 0 waits forever like NCCL).  When one expires the kernel poisons what it
 * could not compute with NaN bit patterns and raises an error word in host-mapped memory; from then on EVERY
 * po_plan_apply* / po_plan_pullback on the context returns PO_ERR_CUDA (po_plan_apply_host reports it for its own apply).
 * This call synchronises the device, returns the word (0 = healthy; 0x1-0xff: a streaming kernel's barrier; 0x100: a peer
 * never acknowledged a staging region; 0x200: a peer's data never arrived) and CLEARS it. */
int32_t po_ctx_device_status(po_ctx* ctx, int32_t* status);

/* ---- plans: A(theta)*x and A'(theta)*y  (src/ParOperator.jl:221-223) ---------------- */
int32_t po_plan_create(po_ctx* ctx, const po_node* nodes, int32_t n_nodes, const int32_t* child_index,
                       int32_t n_child_index, const int64_t* aux, int32_t n_aux, int32_t root, int64_t batch,
                       uint32_t flags, po_plan** out);
/* The same plan from the reference's JSON wire format (src/ASTSerialization.jl:6-58 and the per-type to_Dict methods: ParDFT.jl:34,
 * ParIdentity.jl:24, ParRestriction.jl:45, ParMatrix.jl:62, ParDiagonal.jl:33, ParTensor.jl:119, ParAdjoint.jl:23, ParCompose.jl:100,
 * ParKron.jl:311), extended with the distributed node types its registry leaves out (ParRepartition {T, global_size, dims_in, dims_out,
 * rank}, ParBroadcasted {of, root}, ParDistributed {of, rank, size}).  Parametric leaves get one parameter slot per distinct "id", in
 * order of first appearance (depth first, "of" lists left to right): po_plan_num_params / po_plan_param_id tell the caller which
 * parameter array goes into which slot of `params`. */
int32_t po_plan_create_json(po_ctx* ctx, const char* json, int64_t batch, uint32_t flags, po_plan** out);
int32_t po_plan_num_params(const po_plan* plan, int32_t* n_params);
int32_t po_plan_param_id(const po_plan* plan, int32_t slot, char* buf, size_t buf_bytes);
int32_t po_plan_destroy(po_plan* plan);
int32_t po_plan_domain(const po_plan* plan, int64_t* n, int32_t* dtype);
int32_t po_plan_range(const po_plan* plan, int64_t* n, int32_t* dtype);
int32_t po_plan_workspace_bytes(const po_plan* plan, size_t* bytes);
int32_t po_plan_num_steps(const po_plan* plan, int32_t* n_steps);
/* human readable step list ("dense_dft r2c I=.. n=.. m=.. O=.."), for tests and logs */
int32_t po_plan_describe(const po_plan* plan, char* buf, size_t buf_bytes);
/* kernels the last apply launched (bench.py's gpu_launches is summed from this) */
int32_t po_plan_launch_count(const po_plan* plan, int64_t* launches);
/* algorithmic HBM bytes of one apply: compulsory read of x + write of y (+ parameters) */
int32_t po_plan_algorithmic_bytes(const po_plan* plan, int64_t* bytes);

/* Per-step device timing: while enabled every apply brackets each step with CUDA events on the
 * caller's stream; po_plan_step_profile waits for the recorded work and returns, per step, the
 * summed milliseconds, the number of timed launches and the step's algorithmic bytes (its
 * compulsory input + output + parameter traffic).  bench.py's roofline line comes from this. */
int32_t po_plan_set_profiling(po_plan* plan, int32_t enable);
int32_t po_plan_step_profile(po_plan* plan, double* ms_sum, int64_t* calls, int64_t* bytes, int32_t capacity,
                             int32_t* n_steps);
/* same for the reverse-mode launches of po_plan_pullback (entry i = the cotangent kernels of step i) */
int32_t po_plan_pullback_step_profile(po_plan* plan, double* ms_sum, int64_t* calls, int32_t capacity, int32_t* n_steps);

/* y = A(theta) * x.  x: (Domain, batch) device, y: (Range, batch) device.
 * A plan compiled for batch == 0 (an empty operand, `0 in size(x)` of src/ParDFT.jl:26,30 / src/ParCommon.jl:73) applies, tapes and pulls
 * back as a no-op: PO_OK, nothing launched, x / y / xbar may be NULL and parameter cotangents are left as the caller initialised them.
 * A rank of a distributed operator whose shard is empty on one side (more ranks than rows along a dim) passes NULL for that side; its
 * local steps are skipped and its ParRepartition still takes part in the exchange. */
int32_t po_plan_apply(po_plan* plan, const void* x, void* y, const void* const* params, int32_t n_params,
                      void* workspace, size_t workspace_bytes, void* stream);
/* same, with HOST x/y (pinned or pageable): H2D, apply, D2H on `stream`, then waits for the
 * stream.  This is the reference-facing `A*x` on an `Array` handed to the GPU engine. */
int32_t po_plan_apply_host(po_plan* plan, const void* x_host, void* y_host, const void* const* params,
                           int32_t n_params, void* stream);
/* same without the wait: the three operations are only ENQUEUED on `stream` (x_host / y_host should be page-locked, otherwise the
 * runtime stages the copies synchronously); y_host is valid, and x_host / params may be reused, once the caller has synchronised the
 * stream.  A host pipeline (upload of batch k+1 and download of batch k-1 around the apply of batch k, bench.py's e2e) is built from
 * this call on two or more streams; successive host applies of ONE plan must stay on ONE stream (the plan owns the device copies). */
int32_t po_plan_apply_host_async(po_plan* plan, const void* x_host, void* y_host, const void* const* params,
                                 int32_t n_params, void* stream);

/* ---- reverse mode (Zygote/ChainRules surface of examples/fno.jl) -------- */
/* bytes of the tape that po_plan_apply_taped fills (inputs of parametric steps only) */
int32_t po_plan_tape_bytes(const po_plan* plan, size_t* bytes);
int32_t po_plan_apply_taped(po_plan* plan, const void* x, void* y, const void* const* params, int32_t n_params,
                            void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream);
/* xbar = (dy/dx)^H ybar and param_grads[slot] = cotangent of params[slot] (overwritten; slots
 * that are NULL are skipped).  Uses TRUE adjoints for the DFTs (AbstractFFTs rules), not the
 * inverse-as-adjoint of src/ParDFT.jl:30-31. */
int32_t po_plan_pullback(po_plan* plan, const void* tape, const void* ybar, void* xbar, const void* const* params,
                         void* const* param_grads, int32_t n_params, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- single-operator conveniences (thin one-node plans, cached per ctx) ------------- */
/* ParDFT / its adjoint along dim 1 of an (n_in, batch) matrix: src/ParDFT.jl:26-32 */
int32_t po_dft(po_ctx* ctx, int32_t in_dtype, int64_t n, int32_t adjoint, int64_t batch, const void* x, void* y,
               void* stream);
/* ParRestriction / adjoint: src/ParRestriction.jl:13-39; ranges = [s1,e1,s2,e2,...] 1-based */
int32_t po_restrict(po_ctx* ctx, int32_t dtype, int64_t n, const int64_t* ranges, int32_t n_ranges, int32_t adjoint,
                    int64_t batch, const void* x, void* y, void* stream);

/* ParRepartition / adjoint as a one-node plan: src/ParRepartition.jl:140-201.  x: (prod(local_in), batch), y: (prod(local_out), batch);
 * needs a communicator of prod(dims_in) ranks on the context. */
int32_t po_repartition(po_ctx* ctx, int32_t dtype, int32_t ndim, const int64_t* global_size, const int32_t* dims_in, const int32_t* dims_out,
                       int32_t rank, int32_t adjoint, int64_t batch, const void* x, void* y, void* stream);
/* (A_1 (x) ... (x) A_N) * x for LEAF factors: src/ParKron.jl:190-220.  `factors[k]` is Kron factor k+1 as a po_node (aux slices index
 * `aux`), `order` the 1-based application order of src/ParKron.jl:33-57; plan and workspace are cached on the context. */
int32_t po_kron_apply(po_ctx* ctx, const po_node* factors, int32_t n_factors, const int64_t* aux, int32_t n_aux, const int32_t* order,
                      int64_t batch, const void* x, void* y, const void* const* params, int32_t n_params, void* stream);

/* y = gelu.(x1 .+ x2) in one pass (the block epilogue of examples/fno.jl:40-44); real dtypes;
 * x2 may be NULL; y may alias x1. */
int32_t po_add_gelu(po_ctx* ctx, int32_t dtype, int64_t n, const void* x1, const void* x2, void* y, void* stream);

/* Fused block epilogue of examples/fno.jl:34-44 in ONE pass:
 *     y = act.( s .+ (I_grid (x) W) x .+ b ),   act = gelu when apply_gelu != 0, identity otherwise
 * W: c_out x c_in column-major ParMatrix parameter (src/ParMatrix.jl:42-46); x: (c_in, npts); s: (c_out, npts) or NULL
 * (the spectral-convolution output S x); b: c_out bias or NULL (src/ParMatrix.jl:53-58); y: (c_out, npts).
 * Channels are the fastest dim (the ParKron(ParIdentity(grid), ParMatrix) layout of src/ParKron.jl:190-220).  Real dtypes. */
int32_t po_block_epilogue(po_ctx* ctx, int32_t dtype, int64_t c_in, int64_t c_out, int64_t npts, const void* W, const void* x,
                          const void* s, const void* b, int32_t apply_gelu, void* y, void* stream);

/* Reverse mode of po_block_epilogue (what Zygote does to examples/fno.jl:40-44): with z = s + (I (x) W) x + b recomputed (nothing taped),
 *     g = act'(z) .* ybar,   sbar = g,   xbar = (I (x) W') g,   Wbar = sum_pts g x',   bbar = sum_pts g.
 * sbar: (c_out, npts), REQUIRED (it is also the workspace); xbar (c_in, npts), Wbar (c_out x c_in column-major), bbar (c_out) may be NULL. */
int32_t po_block_epilogue_pullback(po_ctx* ctx, int32_t dtype, int64_t c_in, int64_t c_out, int64_t npts, const void* W, const void* x,
                                   const void* s, const void* b, int32_t apply_gelu, const void* ybar, void* sbar, void* xbar, void* Wbar,
                                   void* bbar, void* stream);

/* ---- host integer bookkeeping (bit-exact rows a15-a17 of SURVEY 8a; no GPU needed) --- */
/* src/ParCommon.jl:57-64 */
int64_t po_local_size(int64_t global_size, int64_t rank, int64_t num_ranks);
/* ParRepartition send/recv plan, src/ParRepartition.jl:14-110.  Boxes are written as
 * [peer, s_1, e_1, ..., s_N, e_N] records (1-based inclusive local ranges), first axis of the
 * peer product fastest, into send_out / recv_out (capacity in records); counts returned. */
int32_t po_repartition_plan(int32_t ndim, const int64_t* global_size, const int32_t* dims_in, const int32_t* dims_out,
                            int32_t rank, int64_t* local_in, int64_t* local_out, int64_t* send_out,
                            int32_t send_capacity, int32_t* n_send, int64_t* recv_out, int32_t recv_capacity,
                            int32_t* n_recv);

/* ---- communicator (ParRepartition / ParBroadcasted over NCCL) ------- */
int32_t po_comm_unique_id(void* id_128_bytes);
int32_t po_comm_init_rank(po_ctx* ctx, int32_t n_ranks, int32_t rank, const void* id_128_bytes);
/* Single process driving several devices (one context each, rank = position in `ctxs`): symmetric buffers are mapped with peer
 * access; carries ParRepartition (NVLink peer-store kernel) only -- parameter collectives need one thread per device + po_comm_init_rank. */
int32_t po_comm_init_all(int32_t n_ctx, po_ctx* const* ctxs);
/* Bring-your-own transport (CUDA-aware MPI.jl in a Julia host, gloo in the CPU tests): the
 * library still packs/unpacks with its own kernels and hands the per-peer staging buffers to
 * `exchange` (an all-to-all-v: post every recv, every send, complete all -- the
 * Irecv!/Isend/Waitany loop of src/ParRepartition.jl:162-189) and parameter traffic to
 * `collective` (op 0 = bcast from root, 1 = sum-reduce to root, 2 = sum-allreduce).  Both
 * return 0 on success and must have completed, or be stream-ordered on `stream`, on return. */
typedef int32_t (*po_exchange_fn)(void* user, int32_t n_send, const int32_t* send_peers, const void* const* send_bufs,
                                  const size_t* send_bytes, int32_t n_recv, const int32_t* recv_peers,
                                  void* const* recv_bufs, const size_t* recv_bytes, void* stream);
typedef int32_t (*po_collective_fn)(void* user, int32_t op, const void* send, void* recv, int64_t count,
                                    int32_t dtype, int32_t root, void* stream);
int32_t po_comm_init_external(po_ctx* ctx, int32_t n_ranks, int32_t rank, po_exchange_fn exchange,
                              po_collective_fn collective, void* user);
int32_t po_comm_destroy(po_ctx* ctx);
int32_t po_comm_rank(const po_ctx* ctx, int32_t* rank, int32_t* n_ranks);
/* MPI.bcast of a parameter array from root, src/ParBroadcasted.jl:26-31 */
int32_t po_bcast(po_ctx* ctx, void* buf, size_t bytes, int32_t root, void* stream);
/* MPI.Reduce(SUM) of a parameter gradient to root, on device, src/ParBroadcasted.jl:41-57 */
int32_t po_reduce_sum(po_ctx* ctx, const void* send, void* recv, int64_t count, int32_t dtype, int32_t root,
                      void* stream);
/* MPI.Allreduce(SUM), src/ParReduce.jl:15-40 */
int32_t po_allreduce_sum(po_ctx* ctx, const void* send, void* recv, int64_t count, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAROPS_H */
