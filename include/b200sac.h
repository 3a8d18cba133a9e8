/*
 * b200sac.h -- C ABI of the B200-native SAC learner hot path.
 *
 * The reference (SKSKSK94/Distributed_SAC) is pure Python and has no FFI; the seam
 * this library sits behind is the Python method surface of `Learner` /
 * `ReplayBuffer`.  Each entry point below names the reference code it replaces
 * (paths relative to the reference checkout):
 *
 *   b200sac_step*, b200sac_update
 *                            Learner.update() -> update_SAC()
 *                            LunarLander_Distributed_SAC/src/learner.py:246-264,203-239
 *                            MT1_Distributed_VSAC/src/learner.py:205-244,251-269
 *                            MT10_Distributed_MTSAC/src/learner.py:253-325,332-352
 *                            MT10_Distributed_CARE/src/learner.py:281-404 (care = 1 | 2), MT1_Distributed_CARE/src/learner.py:247-348
 *   b200sac_replay_*         ReplayBuffer.sample()/__len__ and the append in run()
 *                            LunarLander_Distributed_SAC/src/replay_buffer.py:43-77
 *                            MT10_Distributed_MTSAC/src/replay_buffers.py:47-107
 *   b200sac_layout/export/import
 *                            build_model/build_optimizer/get_parameters/save_checkpoint
 *                            LunarLander_Distributed_SAC/src/learner.py:100-124,144-163,272-276
 *   b200sac_soft_update      Learner.soft_update()            learner.py:126-137
 *   b200sac_publish_*        Learner.get_parameters()         learner.py:272-276 (called at :298-299)
 *   b200sac_blob_*           server.set('parameters', _pickle.dumps(get_parameters())) in Learner.run()
 *                            learner.py:298-299 (MT10_Distributed_CARE/src/learner.py:442-443): the byte string, device-assembled
 *   b200sac_act              Actor.get_action() for a batch of environments
 *                            LunarLander_Distributed_SAC/src/model.py:67-82; MT10_Distributed_CARE/src/player.py:199-209 (CARE handles)
 *
 * Conventions: plain pointers and sizes only (no torch types).  Every function
 * returns 0 on success or a negative b200sac_status; a message for the calling
 * thread is available from b200sac_last_error().  No exceptions cross the
 * boundary.  All device work is enqueued on the caller's stream (a
 * cudaStream_t passed as void*; NULL = legacy default stream).  A handle is
 * single-caller (not re-entrant); distinct handles are independent.
 * `replicas` independent learners of the same shape can live in one handle and
 * are stepped together (grouped launches); every per-learner array then has a
 * leading replica dimension.
 */
#ifndef B200SAC_H_
#define B200SAC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SAC_MAX_HIDDEN 8

typedef enum b200sac_status {
  B200SAC_OK = 0,
  B200SAC_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  B200SAC_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  B200SAC_ERR_NOMEM = -3,
  B200SAC_ERR_STATE = -4      /* call not valid in this state (e.g. replay too small) */
} b200sac_status;

/* Mirrors the cfg JSON + hard-coded dims of the reference learners
 * (cfg/*.json; LL/learner.py:62-81; MS/learner.py:62-98). */
typedef struct b200sac_cfg {
  int32_t state_dim;                 /* raw state width (8 / 39) */
  int32_t act_dim;                   /* 2 / 4 */
  int32_t num_tasks;                 /* 0 = single task; T>0 = one-hot appended to the state (mtobs) */
  int32_t n_actor_hidden;            /* number of hidden layers */
  int32_t n_critic_hidden;
  int32_t actor_hidden[B200SAC_MAX_HIDDEN];
  int32_t critic_hidden[B200SAC_MAX_HIDDEN];
  int32_t batch;                     /* minibatch per learner */
  int32_t weighted_loss;             /* MTSAC use_weighted_loss (== extra 1/batch) */
  int32_t replicas;                  /* independent learners in this handle (>= 1) */
  int32_t precision;                 /* 0 = fp32 FFMA; 1 = 3xTF32 on tcgen05 for hidden layers */
  int32_t care;                      /* CARE state encoders (MT10_Distributed_CARE/src/{context,state}_encoder.py):
                                        1 = CARE(M) use_modified_care=true: frozen embedding, per-encoder mlp_context;
                                        2 = CARE(O) use_modified_care=false (= MT1_Distributed_CARE): trainable context
                                            encoder z = mlp(header(relu(E[t]))) with its own Adam (lr_ctx) */
  double gamma, tau, reward_scale;
  double lr_actor, lr_critic, lr_alpha;
  double action_scale;               /* k = (hi - lo) / 2 */
  double beta1, beta2, adam_eps;
  double log_alpha_init;
  /* CARE only (cfg "encoder" block): */
  int32_t num_encoders;              /* K (num_encoders) */
  int32_t n_mix_hidden;              /* hidden_dims_mixtureEnc (also the attention trunk's hidden dims) */
  int32_t mix_hidden[B200SAC_MAX_HIDDEN];
  int32_t mix_out;                   /* output_dim_mixtureEnc */
  int32_t ctx_in;                    /* RoBERTa_embedding_dim (768) */
  int32_t n_ctx_hidden;              /* hidden_dims_contextEnc */
  int32_t ctx_hidden[B200SAC_MAX_HIDDEN];
  int32_t ctx_out;                   /* output_dim_contextEnc */
  int32_t emb_dim;                   /* embedding_dim_contextEnc: CARE(O) header = Linear(ctx_in, 2e), Linear(2e, e) */
  double tau_se;                     /* state_encoder_tau */
  double lr_ctx;                     /* lr_contextEnc (CARE(O) context-encoder Adam) */
} b200sac_cfg;

typedef struct b200sac_tensor_desc {
  char name[48];                     /* canonical name, e.g. "q1_target.2.weight" */
  int64_t offset;                    /* in floats, inside one replica's parameter arena */
  int32_t rows, cols;                /* weight: (out, in) row-major; bias/log_alpha: (n, 1) */
  int32_t trainable;                 /* 1 -> has Adam m/v at the same offset in the m/v arenas */
  int32_t opt;                       /* 0 critic, 1 actor, 2 alpha, 3 context encoder, -1 none (targets, embedding) */
  int32_t pitch;                     /* floats between consecutive rows (>= cols): first-layer weights are padded to a
                                        multiple of 4 so that TMA can address them; pad columns are and stay zero */
  int32_t reserved;
} b200sac_tensor_desc;

typedef struct b200sac b200sac_t;               /* learner handle */
typedef struct b200sac_replay b200sac_replay_t; /* replay ring handle */

/* arenas addressable through export/import */
enum { B200SAC_PARAMS = 0, B200SAC_ADAM_M = 1, B200SAC_ADAM_V = 2, B200SAC_GRADS = 3 };

const char* b200sac_last_error(void);
const char* b200sac_version(void);

/* Pure host function (no GPU needed): parameter layout for a cfg.
 * Writes up to `cap` descriptors, returns the total count in *n and the arena
 * size (floats per replica; trainable prefix length) in *arena_floats / *trainable_floats. */
int b200sac_layout(const b200sac_cfg* cfg, b200sac_tensor_desc* out, int32_t cap, int32_t* n,
                   int64_t* arena_floats, int64_t* trainable_floats);

/* Create a learner on CUDA device `device`.  Parameters are Xavier-uniform /
 * zero-bias initialised from `seed` (replica r uses seed + r); targets = locals. */
int b200sac_create(const b200sac_cfg* cfg, int32_t device, uint64_t seed, b200sac_t** out);
int b200sac_destroy(b200sac_t* h);

/* Copy one replica's arena out / in.  `buf` may be a host or a device pointer. */
int b200sac_export(b200sac_t* h, int32_t which, int32_t replica, float* buf, int64_t n_floats, void* stream);
int b200sac_import(b200sac_t* h, int32_t which, int32_t replica, const float* buf, int64_t n_floats, void* stream);
/* n_floats floats at `offset` of one replica's arena into HOST memory (the logger's read of log_alpha, learner.py:199:
 * a handful of floats instead of the whole arena).  Synchronises the stream. */
int b200sac_read_range(b200sac_t* h, int32_t which, int32_t replica, int64_t offset, int64_t n_floats, float* out_host,
                       void* stream);
/* Device address of an arena (replica 0; replicas are contiguous, stride = arena floats).  For
 * zero-copy views and for the one-time NCCL broadcast of initial weights done by the host shim. */
int b200sac_arena_ptr(b200sac_t* h, int32_t which, float** dev_ptr, int64_t* floats_per_replica);
/* Adam step counters {critic, actor, alpha, context encoder (CARE(O), else unused)} of one replica. */
int b200sac_get_steps(b200sac_t* h, int32_t replica, int64_t steps[4]);
int b200sac_set_steps(b200sac_t* h, int32_t replica, const int64_t steps[4]);

/* One gradient step from caller-provided DEVICE minibatches, each
 * [replicas][batch][width] fp32: s,s2 width obs(=state_dim+num_tasks), a width act, r,d width 1.
 * eps_next / eps_cur: [replicas][batch][act] standard-normal noise for the two
 * rsample() calls (next-state actor first), or NULL to draw in-kernel (Philox).
 * Losses land in an internal device ring; fetch with b200sac_read_losses. */
int b200sac_step(b200sac_t* h, const float* s, const float* a, const float* r, const float* s2,
                 const float* d, const float* eps_next, const float* eps_cur, void* stream);

/* Same, from HOST buffers (any host memory; staged through the handle's pinned
 * double buffer with cudaMemcpyAsync on a side stream).  out_losses (host,
 * [replicas][4] = critic, actor, alpha-loss, entropy) may be NULL; if given the call
 * synchronises and fills it -- this is the reference's `update()` contract. */
int b200sac_step_host(b200sac_t* h, const float* s, const float* a, const float* r, const float* s2,
                      const float* d, const float* eps_next, const float* eps_cur, float* out_losses,
                      void* stream);

/* n_steps gradient steps, each sampling its own minibatch from the replay ring
 * (uniform without replacement; per-task B/T when num_tasks > 0) with in-kernel noise.
 * Device-resident ring: everything is enqueued asynchronously (one CUDA graph per step).
 * Pinned-host ring: indices drawn on the host, rows gathered into pinned staging and
 * copied on a side stream, overlapped with the previous step. */
int b200sac_step_sampled(b200sac_t* h, b200sac_replay_t* rb, int32_t n_steps, void* stream);

/* Capture and instantiate every CUDA graph b200sac_step_sampled / b200sac_update can launch for this ring (device ring:
 * the 8-, 4-, 2- and 1-step graphs a run is cut into; host ring: the staged one-step graph of both staging slots) without
 * running a step, so that no later call pays graph construction (milliseconds).  Synchronises the stream. */
int b200sac_prepare(b200sac_t* h, b200sac_replay_t* rb, void* stream);

/* Learner.update() (learner.py:246-264) in one call: b200sac_step_sampled(h, rb, 1) followed by the losses of that
 * step, out[replicas][4] = {critic, actor, alpha loss, entropy}.  Synchronises the stream. */
int b200sac_update(b200sac_t* h, b200sac_replay_t* rb, float* losses_host, void* stream);

/* Losses of the most recent `n_last` steps (<= 1024 kept): out[n_last][replicas][4].
 * Synchronises the stream. */
int b200sac_read_losses(b200sac_t* h, int32_t n_last, float* out_host, void* stream);

/* Polyak update of the target critics outside a step (Learner.soft_update; tau = 1 is
 * the hard copy Learner.run() does before training, learner.py:287-288). */
int b200sac_soft_update(b200sac_t* h, double tau, void* stream);

/* Batched policy inference for n <= 2*batch observation rows obs[n][obs_dim] (obs_dim = state_dim + num_tasks), host or
 * device memory: actions[n][act_dim] = k * tanh(mu + std * eps) -- Actor.get_action
 * (LunarLander_Distributed_SAC/src/model.py:67-82, MT10_Distributed_MTSAC/src/model.py:58-73) vectorised over environments.
 * CARE handles (n <= batch): the rows are first encoded like the player does it, z = context_encoder(mtobs) then the
 * actor's state encoder -- the hard copy of the critic's -- (MT10_Distributed_CARE/src/player.py:199-209, model.py:90-114).
 * eps (nullable, [n][act_dim]) injects the noise; eps == NULL: stochastic = 1 draws fresh Philox noise; stochastic = 0
 * returns the deterministic action k * tanh(mu) of the VSAC / MTSAC / CARE actors (MT1_Distributed_VSAC/src/model.py:63,80);
 * stochastic = -1 returns k * mu, no squashing, which is what the LunarLander actor's get_action(stochastic=False) does
 * (LunarLander_Distributed_SAC/src/model.py:67-82).  Synchronises the stream. */
int b200sac_act(b200sac_t* h, int32_t replica, int32_t n, const float* obs, const float* eps, int32_t stochastic,
                float* actions_out, void* stream);

/* Publication path = Learner.get_parameters() (LunarLander_Distributed_SAC/src/learner.py:272-276,298-299;
 * MT10_Distributed_CARE/src/learner.py:412-417,442-443), which the reference runs after every update.
 * publish_begin enqueues, in stream order, a consistent device snapshot of `n_ranges` ranges
 * [offsets[i], offsets[i]+counts[i]) of replica `replica`'s parameter arena (offsets from b200sac_layout)
 * and starts its copy to pinned host memory on a private stream; it does not block, and steps enqueued
 * afterwards overlap the copy.  publish_wait blocks until the OLDEST uncollected snapshot has landed and returns its
 * packed ranges.  Two snapshot slots: publish_begin may be called again before the previous snapshot is collected (the
 * caller then turns snapshot k into the published blob while step k+1 runs and snapshot k+1 lands); the pointer stays
 * valid until the second publish_begin after the one that produced it.  A third begin without a wait supersedes the oldest. */
int b200sac_publish_begin(b200sac_t* h, int32_t replica, int32_t n_ranges, const int64_t* offsets, const int64_t* counts,
                          void* stream);
int b200sac_publish_wait(b200sac_t* h, const float** host_ptr, int64_t* n_floats);

/* Blob publication: the byte string Learner.run() stores under the Redis key 'parameters' after every update
 * (LunarLander_Distributed_SAC/src/learner.py:298-299: _pickle.dumps({'actor': state_dict}); read back by
 * Player.pull_parameters, player.py:75-85) assembled on the device.  blob_template: `image` = the constant bytes of the
 * string (any values in the payload positions); payload float j (j < n_floats) is the 4 bytes at dst_byte[j] and holds
 * element src_index[j] of replica `replica`'s parameter arena (b200sac_layout offsets; pitch padding / transposed views
 * are expressed through the index map).  blob_begin enqueues, in stream order, the gather of all payload floats into a
 * device copy of the image and starts ONE D2H copy of it on a private stream; it does not block and later steps overlap
 * the copy.  blob_wait blocks until the OLDEST uncollected blob has landed in pinned memory and returns it (valid until
 * the second blob_begin after the one that produced it; two image slots, as for publish_begin/wait). */
int b200sac_blob_template(b200sac_t* h, int32_t replica, const uint8_t* image, int64_t image_bytes, int64_t n_floats,
                          const int32_t* src_index, const int32_t* dst_byte);
int b200sac_blob_begin(b200sac_t* h, void* stream);
int b200sac_blob_wait(b200sac_t* h, const uint8_t** host_ptr, int64_t* n_bytes);

/* Debug / parity access to per-step intermediates of replica `replica`:
 * name in {"y","q1","q2","a_next","logp_next","a_cur","logp_cur","qmin","d_action","d_head","r","d"}, "dq_pi" ([2][B]: d(actor loss)/dQ1, dQ2 -- which twin
 * min(Q1,Q2) routed the gradient to), "q_pi" ([2][B]: Q1, Q2 at (s, a~); layer-chained plan only), "psave" ([2B][act][8]:
 * what the policy head saved per (row, action) -- std, u-mu, tanh(u), action, Jacobian term, the NOISE it used, clamp mask,
 * log-prob term), or a hidden
 * activation whose sign pattern is the ReLU mask the step used: "hA.<l>" [2B][H] (rows [s';s]), "hQ.<l>" / "hP.<l>" /
 * "hT.<l>" [2][B][H] (critic-update pass / actor pass / target pass; the layer-chained plan does not keep hT),
 * "mixH.<inst>.<l>" [K][rows][pitch4(H)] (CARE mixture encoders). */
int b200sac_debug_read(b200sac_t* h, const char* name, int32_t replica, float* out_host, int64_t cap_floats,
                       int64_t* n_floats, void* stream);

/* Eager (no graph) run of `iters` sampled steps from a DEVICE ring with a CUDA event between all
 * launches: out_ms[i] = mean device time of launch i of the step (0 = sampling + ingest).
 * names (nullable): ';'-separated kernel labels.  Measurement helper for bench.py's roofline. */
int b200sac_profile_step(b200sac_t* h, b200sac_replay_t* rb, int32_t iters, float* out_ms, int32_t cap,
                         int32_t* n_out, char* names, int32_t names_cap, void* stream);

/* In-graph timeline: `iters` sampled steps from a DEVICE ring as normal graph launches, with every
 * kernel stamping %globaltimer at entry; out_us[i] = mean start-to-start time of launch i
 * (0 = index sampling, 1 = ingest, then the step's launches in order, as b200sac_profile_step names them). */
int b200sac_graph_timeline(b200sac_t* h, b200sac_replay_t* rb, int32_t iters, float* out_us, int32_t cap,
                           int32_t* n_out, void* stream);

/* Stand-alone run of the tcgen05 3xTF32 GEMM kernel on DEVICE arrays (parity tests of the tensor-core
 * path): mode 0 FWD C[M][N] = act(A[M][K] B[N][K]^T + bias), 1 DGRAD C[M][N] = (A[M][K] B[K][N]) * [mask > 0],
 * 2 WGRAD C[M][N] = A[K][M]^T B[K][N], C2[M] = column sums of A.  Synchronises. */
int b200sac_tc_gemm_test(int32_t mode, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* B,
                         int32_t ldb, const float* bias, const float* mask, int32_t ldmask, float* C, int32_t ldc,
                         float* C2, int32_t relu, void* stream);

/* The same for any GEMM engine of the step: engine 0 = 32x32-tile fp32 FFMA, 1 = tcgen05 3xTF32,
 * 2 = thin backward kernel (modes 1/2 with N <= 16: the input-layer weight / input gradients). */
int b200sac_gemm_test(int32_t engine, int32_t mode, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda,
                      const float* B, int32_t ldb, const float* bias, const float* mask, int32_t ldmask, float* C,
                      int32_t ldc, float* C2, int32_t relu, void* stream);

/* Number of kernels one step launches (for bench.py's gpu_launches accounting). */
int b200sac_launches_per_step(b200sac_t* h, int32_t* n);

/* ---- replay ring ---------------------------------------------------------------- */
/* where: 0 = device-resident ring (HBM), 1 = pinned host DRAM ring.
 * capacity = transitions per replica (per task: capacity / num_tasks when num_tasks > 0). */
int b200sac_replay_create(b200sac_t* h, int64_t capacity, int32_t where, uint64_t seed, b200sac_replay_t** out);
int b200sac_replay_destroy(b200sac_replay_t* rb);
/* Append n transitions (HOST arrays, fp32, same widths as b200sac_step) to one replica's ring.
 * Thread-safe against one concurrent sampler. */
int b200sac_replay_push(b200sac_replay_t* rb, int32_t replica, int64_t n, const float* s, const float* a,
                        const float* r, const float* s2, const float* d);
/* Fill every replica's ring with `n` synthetic transitions generated on the device
 * (s,s'~N(0,1), a~U(-1,1), r~N(0,1), d~Bernoulli(0.01), task = i mod T). Benchmark helper. */
int b200sac_replay_fill_synthetic(b200sac_replay_t* rb, int64_t n, uint64_t seed, void* stream);
/* min over tasks of the per-task fill, like ReplayBuffer.__len__ in the MT variant. */
int b200sac_replay_size(b200sac_replay_t* rb, int32_t replica, int64_t* n);
/* Draw one minibatch the way step_sampled would and return it in HOST arrays
 * (reference-shaped ReplayBuffer.sample()); idx_out (nullable) gets the ring indices. */
int b200sac_replay_sample(b200sac_replay_t* rb, int32_t replica, float* s, float* a, float* r, float* s2,
                          float* d, int64_t* idx_out);

#ifdef __cplusplus
}
#endif
#endif /* B200SAC_H_ */
