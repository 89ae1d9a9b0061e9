/*
 * b200dock.h -- C ABI of libb200dock.so: the B200-native DiffBindFR reverse-diffusion hot path.
 *
 * Plain C: raw device/host pointers, sizes, a cudaStream_t passed as void*, int status codes.
 * No torch types, no exceptions across the boundary.  One handle per (process, device); a handle
 * is not thread-safe (the reference drives one nn.Module from one Python thread,
 * druglib/datasets/builder.py:177-183).  The caller owns every input/output buffer, the library owns only
 * the workspace inside the handle (grown lazily, freed by b200dock_destroy).
 *
 * Synchronisation, per entry point:
 *   b200dock_create / destroy / load_weights / mdn_load_* : synchronous (device-wide copies).
 *   b200dock_score / b200dock_sample : all kernels are enqueued on the caller's stream (plus one internal side stream
 *       joined back before return); the call then ENDS WITH ONE cudaStreamSynchronize that reads back the capacity flag
 *       and the edge counts (B200_ERR_CAPACITY is reported here).  b200dock_set_deferred_check(h, 1) removes that
 *       synchronisation: the calls return with the work merely enqueued and b200dock_check(h, stream) delivers the
 *       verdict whenever the caller chooses to synchronise.  A family whose edge list overflows is emptied ON THE DEVICE
 *       (no out-of-bounds access); the results of that evaluation are then meaningless and the status says so.
 *   b200dock_sample_host : synchronous (H2D, sampling, D2H, one stream synchronisation).
 *   b200dock_mdn_encode / mdn_score : asynchronous on the caller's stream, no synchronisation.
 *   b200dock_debug_tap, b200dock_tp_kernel_time_ms : synchronous (tests / benchmarks only).
 *
 * Reference interfaces replaced (paths under /root/reference):
 *   b200dock_score   <- TensorProductModel.forward(data) -> (tr, rot, tor, sc_tor)
 *                       druglib/models/Docking/interaction/tpscore.py:462-573
 *   b200dock_sample  <- DiffBindFR.sample(data, visualize)   druglib/models/Docking/scFlex.py:124-250
 *                       (score net + SDE perturbation :154-205 + update_batchlig_pos
 *                        druglib/utils/bio_utils/conformer_utils.py:420-473 + side-chain rebuild
 *                        druglib/utils/obj/prot_math.py:243-291)
 *   b200dock_load_weights <- load_checkpoint(strict=True) state_dict contract
 *                       druglib/core/runner/checkpoint.py:403-459 (packed by diffbindfr_b200/packer.py)
 *   b200dock_sample_host  same as b200dock_sample with HOST buffers: pinned staging + H2D/D2H inside
 *                       (what DiffBindFR.forward_test + MDLDataParallel.scatter do, scFlex.py:66-81,
 *                        druglib/core/runner/parallel/_functions.py:75-94)
 */
#ifndef B200DOCK_H_
#define B200DOCK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID 1      /* bad argument / inconsistent sizes */
#define B200_ERR_CUDA 2         /* a CUDA runtime call failed; see b200dock_last_error */
#define B200_ERR_CAPACITY 3     /* an edge list overflowed its workspace capacity */
#define B200_ERR_STATE 4        /* weights / plans not loaded */

#define B200_MAX_PATHS 16
#define B200_MAX_BLOCKS 4
#define B200_H_STRIDE 168       /* node feature row stride (floats) */
#define B200_K_PAD 160          /* padded K of the per-edge weight generator (144 + bias row + pad) */

typedef struct B200Handle B200Handle;

/* One tensor-product path (e3nn 'uvw' instruction, mul2 == 1), see diffbindfr_b200/spec.py:Path. */
typedef struct {
  int32_t l1, l2, lo;
  int32_t U, Wd;
  int32_t in1_off, in2_off, out_off;
  int32_t z_off;            /* offset of this path's [U][2lo+1] block in the per-edge Z vector */
  int32_t cg_off, cg_n;     /* range in the sparse Clebsch-Gordan table of the plan */
  int32_t col_off;          /* first column (row of the packed W2) of this path */
} B200Path;

/* One irreps block of a message row, for the equivariant LayerNorm (tpscore.py:20-107). */
typedef struct {
  int32_t off, mul, dim;    /* column offset, multiplicity, 2l+1 */
  int32_t irr_off;          /* offset into mean_shift / affine_weight */
  int32_t bias_off;         /* offset into affine_bias, or -1 when the block is not 0e */
} B200Block;

/* Static description of one TensorProductConvLayer flavour (conv layers 0,1,2,3+ and the torsion conv). */
typedef struct {
  int32_t n_paths;
  B200Path paths[B200_MAX_PATHS];
  int32_t in_dim, sh_dim, out_dim, z_numel, n_cols;  /* n_cols = weight_numel */
  int32_t n_blocks;
  B200Block blocks[B200_MAX_BLOCKS];
  int32_t n_cg;             /* sparse CG entries: (i, j, k) packed as i | j<<8 | k<<16, and values */
  const int32_t* cg_ijk;    /* host pointers, copied by b200dock_create */
  const float* cg_val;
  int32_t n_chunks;         /* column chunks of the packed W2 (each <= 192 columns, one path each) */
  const int32_t* chunk_col; /* [n_chunks] first column */
  const int32_t* chunk_n;   /* [n_chunks] number of columns */
  const int32_t* chunk_path;/* [n_chunks] path index */
} B200ConvPlan;

#define B200_PLAN_L0 0
#define B200_PLAN_L1 1
#define B200_PLAN_L2 2
#define B200_PLAN_L3 3      /* layers 3..5 */
#define B200_PLAN_TOR 4     /* tor_bond_conv / sc_tor_bond_conv */
#define B200_PLAN_FINAL 5   /* final_conv (centre conv) */
#define B200_N_PLANS 6

typedef struct {
  B200ConvPlan plans[B200_N_PLANS];
  int32_t conv_kernel;      /* 0 exact fp32 SIMT (cross-check), 5 fused tcgen05 conv with FP16 hi/lo MMAs + per-row scaling
                               (fp32-grade; message rows + separate scatter), 6 the same on CTA pairs (cta_group::2) with the
                               scatter fused into the epilogue (bit-identical to 5), 11 kernel 6 plus a warpgroup that gathers / converts the next tile's
                               edge input one tile ahead and converts H1 (default of the Python layer; bit-identical) */
  int32_t reserved[7];
  const int32_t* atom14_group;  /* [21][14] restype_atom14_to_rigid_group (protein_constants.py:1177-1199) */
  /* sparse CG tables of the pseudo-torque product harmonics: triples (2,2,0), (1,2,1), (2,2,1) */
  const int32_t* tor_cg_ijk;
  const float* tor_cg_val;
  int32_t tor_cg_off[4];
} B200Config;

/* Weight blob sections (offsets in floats into the blob handed to b200dock_load_weights).
 * Layouts are documented in diffbindfr_b200/packer.py. */
enum {
  B200_W_LIG_NODE = 0,      /* W0t[59][48] b0[48] W3t[48][48] b3[48] */
  B200_W_LIG_EDGE,          /* W0t[74][48] b0 W3t b3   (input order: bond feat 10 | sigma 32 | rbf 32) */
  B200_W_ATOM_EMB,          /* tables [37+22+4+21+2][48] then scalar_lin Wt[80][48] (x_emb 48 | sigma 32) */
  B200_W_ATOM_EDGE,         /* W0t[64][48] b0 W3t b3   (sigma 32 | rbf 32) */
  B200_W_LA_EDGE,
  B200_W_CENTER_EDGE,
  B200_W_TOR_EDGE,          /* W0t[32][48] b0 W3t b3 */
  B200_W_SC_EDGE,
  B200_W_FINAL_FC,          /* W1t[96][96] b1[96] W2t[96][336] b2[336] (alpha folded) */
  B200_W_FINAL_LN,          /* mean_shift[4] affine_weight[4] */
  B200_W_TR_FINAL,          /* W0t[33][48] b0[48] w3[48] b3[1] */
  B200_W_ROT_FINAL,
  B200_W_TOR_FINAL,         /* W0t[96][48] w3[48] */
  B200_W_SC_FINAL,
  B200_W_CONV0,             /* first of 26 conv records: lig[0..5] atom[0..5] al[0..5] la[0..5] tor sc;
                               each: W1t[144][144] b1[144] W2p[n_cols][160] ln_shift[nirr] ln_w[nirr] ln_b[nscalar] */
  B200_W_N_SECTIONS = B200_W_CONV0 + 26
};

/* Collated batch, int32 indices, device pointers for b200dock_score/sample,
 * host pointers for b200dock_sample_host.  Derived index arrays are prepared by the host
 * wrapper (diffbindfr_b200/batch.py). */
typedef struct {
  int32_t B, N_l, N_a, N_r, E_b, n_tor, n_sc;
  int32_t max_lig_atoms;          /* largest ligand of the batch (<= 256) */
  int64_t rot_mask_bytes;         /* total bytes of rot_mask */
  int64_t cross_pairs;            /* sum_g n_lig(g) * n_atom(g): exact bound of the cross edge list */
  int64_t atom_pairs;             /* sum_g n_atom(g)^2: bound of the pocket edge list */
  /* ligand */
  const float* lig_node;          /* [N_l][27] */
  float* lig_pos;                 /* [N_l][3]   in/out (updated in place by sample) */
  const int32_t* lig_ptr;         /* [B+1] */
  const int32_t* lig_batch;       /* [N_l] */
  const int32_t* bond_ptr;        /* [N_l+1] CSR of lig_edge_index by edge_index[0] (stable) */
  const int32_t* bond_dst;        /* [E_b] edge_index[1] in CSR order */
  const int32_t* bond_eid;        /* [E_b] original bond id (row of lig_edge_feat) */
  const float* lig_edge_feat;     /* [E_b][10] */
  const int32_t* tor_bonds;       /* [n_tor][2] lig_edge_index[:, tor_edge_mask] (global atom ids u, v) */
  const int32_t* tor_ptr;         /* [B+1] torsion bonds per graph */
  const uint8_t* rot_mask;        /* concatenated rot_node_mask rows: bond t -> [n_l(graph)] bytes */
  const int64_t* rot_mask_off;    /* [n_tor] byte offset of row t */
  /* pocket */
  const int32_t* pocket_feat;     /* [N_a][5] categorical codes */
  float* rec_atm_pos;             /* [N_a][3]  in/out */
  const int32_t* atom_ptr;        /* [B+1] */
  const int32_t* atom_batch;      /* [N_a] */
  const int32_t* atom_slot;       /* [N_a] residue*14 + atom14 slot of every pocket atom */
  const int32_t* res_ptr;         /* [B+1] */
  const uint8_t* atom14_mask;     /* [N_r][14] */
  const int32_t* sequence;        /* [N_r] */
  const float* backbone_transl;   /* [N_r][3] */
  const float* backbone_rots;     /* [N_r][9] */
  const float* default_frame;     /* [N_r][8][16] */
  const float* rigid_group_pos;   /* [N_r][14][3] */
  float* torsion_angle;           /* [N_r][5]  in/out */
  const int32_t* sc_bonds;        /* [n_sc][2] torsion_edge_index[sc_torsion_edge_mask] (atom ids j, k) */
  const int32_t* sc_index;        /* [N_r][4] rank of (r, chi) among masked entries, or -1 */
  const int32_t* sc_ptr;          /* [B+1] chi bonds per graph (sc_bonds is grouped by graph) */
} B200Batch;

/* Per-evaluation conditioning written by set_time (scFlex.py:104-122); per graph so that
 * TensorProductModel.forward's contract (data.t per graph) is honoured. */
typedef struct {
  const float* time_emb;          /* [B][32] sinusoidal(1000 t)  (time_emb.py:9-26), computed by the host */
  const float* tr_sigma;          /* [B] */
  const float* rot_score_norm;    /* [B] */
  const float* tor_score_norm2;   /* [n_tor] */
  const float* sc_tor_score_norm2;/* [n_sc]  (already gathered by sc_torsion_edge_mask) */
} B200Cond;

/* Scalars of one reverse-SDE step (scFlex.py:146-205), fp32 as the reference evaluates them. */
typedef struct {
  float t, dt;
  float tr_sigma, rot_score_norm, tor_score_norm2, sc_tor_score_norm2;
  float tr_g2, tr_gs;             /* g^2 and g*sqrt(dt) (0 noise scale handled by zero z) */
  float rot_g2, rot_gs;
  float tor_g2, tor_gs;
  float sc_g2, sc_gs;
  int32_t ode;                    /* 1: perturb = 0.5 g^2 score dt (scFlex.py:162-165,199-200) */
  int32_t reserved;
} B200Step;

int b200dock_create(const B200Config* cfg, int device, B200Handle** out);
void b200dock_destroy(B200Handle* h);
const char* b200dock_last_error(const B200Handle* h);
const char* b200dock_version(void);

/* blob: HOST pointer to n floats; offsets: B200_W_N_SECTIONS float offsets. Synchronous. */
int b200dock_load_weights(B200Handle* h, const float* blob, size_t n, const int64_t* offsets, int n_sections);

/* One score-network evaluation. Outputs: tr[B][3], rot[B][3], tor[n_tor], sc[n_sc] (device). */
int b200dock_score(B200Handle* h, const B200Batch* batch, const B200Cond* cond,
                   float* tr, float* rot, float* tor, float* sc, void* stream);

/* n_steps reverse-SDE steps in place on batch->lig_pos / rec_atm_pos / torsion_angle.
 * noise: device, per step [B*3 | B*3 | n_tor | n_sc] floats (reference draw order scFlex.py:167-204).
 * lig_traj   (optional, device) [n_steps][N_l][3]; atom14_out (device) [N_r][14][3] after the last step;
 * atom14_traj (optional, device) [n_steps][N_r][14][3]. */
int b200dock_sample(B200Handle* h, B200Batch* batch, const B200Step* steps, int n_steps,
                    const float* time_emb /* host [n_steps][32] */, const float* noise,
                    float* lig_traj, float* atom14_out, float* atom14_traj, void* stream);

/* Batch assembly on the device.  `base` holds every COMPLEX once (host pointers, same struct as for
 * b200dock_sample_host); output graph g is a copy of base graph src_graph[g] (index arrays shifted), optionally with a fresh
 * starting pose: LigInit (druglib/datasets/Docking/struct_init.py:16-53: uniform torsions, uniformly random rotation,
 * N(0, tr_sigma_max^2) translation) and SCProtInit (:113-136: chi ~ U(-pi, pi) on the existing chi angles, atom14 rebuilt),
 * drawn from a Philox4x32-10 stream keyed by (seed, stream_id[g]) so that a sample's pose does not depend on the batch or
 * rank it lands in.  Replaces the 40x host collation of the same pocket (druglib/data/collate.py:18-137).  `out` receives
 * DEVICE pointers owned by the handle (valid until the next b200dock_expand_host / destroy); pass it to b200dock_sample.
 * Asynchronous on `stream` after the staging memcpy. */
typedef struct {
  int32_t B_out;
  int32_t randomize;              /* 0: plain replication; 1: LigInit + SCProtInit on the device */
  const int32_t* src_graph;       /* host [B_out] */
  const uint64_t* stream_id;      /* host [B_out] RNG stream of every output graph (e.g. global sample id) */
  uint64_t seed;
  float tr_sigma_max;             /* LigInit(tr_sigma_max), DiffBindFR/configs/diffbindfr_ts.py */
  float reserved;
} B200Expand;
int b200dock_expand_host(B200Handle* h, const B200Batch* base_host, const B200Expand* ex, B200Batch* out_dev, void* stream);

/* Deferred end-of-call check (see "Synchronisation" above). */
int b200dock_set_deferred_check(B200Handle* h, int on);
int b200dock_check(B200Handle* h, void* stream);

/* Same with HOST pointers everywhere (batch arrays, noise, outputs): stages through pinned memory,
 * copies H2D, samples, copies the results D2H and synchronises the stream before returning.
 * lig_out [N_l][3], atom14_out [N_r][14][3]; bytes moved are reported for the e2e measurement. */
int b200dock_sample_host(B200Handle* h, const B200Batch* host_batch, const B200Step* steps, int n_steps,
                         const float* time_emb, const float* noise, float* lig_out, float* atom14_out,
                         uint64_t* h2d_bytes, uint64_t* d2h_bytes, void* stream);

/* MDN scoring head: KarmaDock.scoring(lig_s, lig_pos, pro_s, data, dist_threhold, batch_size)
 * (DiffBindFR/scoring/architecture/KarmaDock_sc.py:87-101, MDN_Block.py:20-79; called from
 * DiffBindFR/common/engines.py:285-294).  lig_s / pro_s come from b200dock_mdn_encode. */
typedef struct {
  int32_t B, N_l, N_r;
  const float* lig_s;             /* [N_l][128] ligand atom embeddings */
  const float* lig_pos;           /* [N_l][3] */
  const int32_t* lig_ptr;         /* [B+1] */
  const float* pro_s;             /* [N_r][128] residue embeddings */
  const float* xyz_full;          /* [N_r][14][3] atom14 coordinates (missing atoms = 0, as in the reference) */
  const int32_t* res_ptr;         /* [B+1] */
} B200MdnBatch;
/* blob (host, 128*128*2 + 128 + 128*30 + 30 floats): Wl_t[128][128] Wr_t[128][128] bias[128] (Linear(256->128) split
 * by input half, BatchNorm(eval) folded), W30_t[128][30] (pi|sigma|mu), b30[30]. Synchronous. */
int b200dock_mdn_load_weights(B200Handle* h, const float* blob, size_t n);
/* score: device [B]. */
int b200dock_mdn_score(B200Handle* h, const B200MdnBatch* batch, float dist_threshold, float* score, void* stream);

/* MDN scorer encoders: KarmaDock.encoding(data) -> (pro_node_s, lig_node_s)
 * (DiffBindFR/scoring/architecture/KarmaDock_sc.py:71-85; GVP_Block.py:63-79 GVP_embedding.forward,
 * GraphTransformer_Block.py:413-424 GraghTransformer.forward; called from DiffBindFR/common/engines.py:285).
 * All pointers are device pointers owned by the caller; indices are int32.  The ligand edges are the covalent subset
 * (the reference indexes edge_s / edge_index with cov_edge_mask before the encoder).  (perm, ptr) is the CSR of the edges
 * by aggregation target in stable edge order: ligand target = col (edge_index[1]), pocket target = dst (edge_index[1]). */
typedef struct {
  int32_t N_r, E_p, N_l, E_l;
  const float* pro_node_s;        /* [N_r][9] */
  const float* pro_node_v;        /* [N_r][3][3] */
  const int32_t* pro_seq;         /* [N_r] residue type, 0..30 */
  const int32_t* pro_src;         /* [E_p] edge_index[0] (message source j) */
  const int32_t* pro_dst;         /* [E_p] edge_index[1] (target i) */
  const float* pro_edge_s;        /* [E_p][21] */
  const float* pro_edge_v;        /* [E_p][1][3] */
  const int32_t* pro_perm;        /* [E_p] */
  const int32_t* pro_ptr;         /* [N_r+1] */
  const float* lig_node_s;        /* [N_l][89] */
  const float* lig_edge_s;        /* [E_l][20] */
  const int32_t* lig_row;         /* [E_l] edge_index[0] */
  const int32_t* lig_col;         /* [E_l] edge_index[1] */
  const int32_t* lig_perm;        /* [E_l] */
  const int32_t* lig_ptr;         /* [N_l+1] */
} B200MdnGraph;
/* Encoder weights: one fp32 blob + B200_MDN_ENC_SECTIONS offsets (in floats, -1 = section absent), packed by
 * diffbindfr_b200/mdn.py::pack_encoder_weights.  Linear weights are stored transposed [in][out]; BatchNorm1d(eval)
 * is folded into the Linear that follows it.  Section order:
 *   0..3    graph transformer node_encoder W,b; edge_encoder W,b
 *   4+14l+j layer l<6: QKV W[128][384],b; edge proj W,b; O_node W,b; node MLP.0 W[128][256],b; node MLP.3 W[256][128];
 *           O_edge W,b; edge MLP.0 W,b; edge MLP.3 W   (edge sections absent in the final layer)
 *   88      W_s embedding [31][31];  89,90 W_v LayerNorm w,b;  91..94 W_v GVP (wh[h][vi], ws_t[si+h][so], ws_b, wv[vo][h])
 *   95,96   W_e LayerNorm;  97..100 W_e GVP
 *   101+24l layer l<3: message GVP 0,1,2 (4 each), norm.0 (2), ff GVP 0,1 (4 each), norm.1 (2)
 *   173,174 W_out LayerNorm;  175..178 W_out GVP (wv absent) */
#define B200_MDN_ENC_SECTIONS 179
int b200dock_mdn_load_encoder_weights(B200Handle* h, const float* blob, size_t n, const int64_t* offsets, int n_offsets);
/* pro_s: device [N_r][128]; lig_s: device [N_l][128]. Asynchronous on `stream`. */
int b200dock_mdn_encode(B200Handle* h, const B200MdnGraph* g, float* pro_s, float* lig_s, void* stream);

/* MDN protein featurisation from the sampler's atom14 output: replaces get_protein_feature's geometry part
 * DiffBindFR/scoring/dataset/protein_feature.py:170-217 and its torch_cluster.knn_graph(CA, k=topk) call.  One graph per pose;
 * edges are emitted grouped by centre (edge_index[1]) in ascending centre order, neighbours by ascending distance, so
 * node_ptr is the CSR b200dock_mdn_encode needs (perm = identity).  All pointers are device pointers. */
typedef struct {
  int32_t B, N_r, topk, max_res;  /* max_res = largest graph (residues); topk <= 32 */
  int64_t E;                      /* sum_g n_g * min(topk, n_g - 1) */
  const int32_t* res_ptr;         /* [B+1] */
  const int64_t* edge_ptr;        /* [B+1] prefix sum of n_g * min(topk, n_g - 1) */
  const float* atom14;            /* [N_r][14][3] (missing atoms = 0) */
  const uint8_t* atom14_mask;     /* [N_r][14] */
  const float* bb_sincos;         /* [N_r][6] backbone dihedral sin/cos (pose independent) */
  float* node_s;                  /* out [N_r][9] */
  float* node_v;                  /* out [N_r][3][3] */
  int32_t* edge_src;              /* out [E] edge_index[0] (neighbour) */
  int32_t* edge_dst;              /* out [E] edge_index[1] (centre) */
  float* edge_s;                  /* out [E][21] */
  float* edge_v;                  /* out [E][1][3] */
  int32_t* node_ptr;              /* out [N_r+1] */
} B200MdnFeat;
int b200dock_mdn_featurize(B200Handle* h, const B200MdnFeat* f, void* stream);

/* Error correction of docked poses: replaces the per-pose `smina.static --minimize` subprocess of
 * druglib/ops/smina/__init__.py:113-146 (smina_min_inplace; called by error_corrector, DiffBindFR/common/engines.py:304-322)
 * and its `--score_only` variant.  One call handles all poses of ONE complex: AutoDock Vina 1.1.2 / smina default scoring function
 * (what the binary evaluates) between the ligand heavy atoms and each pose's own pocket heavy atoms, plus the ligand's
 * intramolecular pairs; mode 1 minimises it by BFGS over translation, rotation (about atom `root`) and torsion increments.
 * Atom typing is an input (X-Score radius + flag bits 1 hydrophobe, 2 donor, 4 acceptor; diffbindfr_b200/vina_types.py).
 * fp64 arithmetic, deterministic.  All pointers are DEVICE pointers; asynchronous on `stream`.
 * Limits: n_lig <= 128, n_tors <= 58, n_rec <= 10000. */
typedef struct {
  int32_t n_pose, n_lig, n_rec, n_tors, root, max_steps, mode, reserved;   /* mode 0 score only, 1 minimise */
  int64_t rec_pose_stride;       /* atoms between consecutive poses in rec_xyz; 0 = one rigid receptor for all poses */
  const float* lig_xyz;          /* [n_pose][n_lig][3] */
  const float* lig_radius; const uint8_t* lig_flags;      /* [n_lig] */
  const float* rec_xyz;          /* [n_pose | 1][n_rec][3] */
  const float* rec_radius; const uint8_t* rec_flags;      /* [n_rec] */
  const int32_t* tors_axis;      /* [n_tors][2]: (atom on the root side, atom that moves); parents before children */
  const uint8_t* tors_mask;      /* [n_tors][n_lig]: 1 = atom moves with this torsion */
  const int32_t* pair_ptr;       /* [n_lig+1] CSR of the intramolecular pair list, every pair listed from both ends */
  const int32_t* pair_idx;       /* [2 n_pairs] */
  double n_rot;                  /* rotor count of the affinity normalisation 1 / (1 + 0.05846 n_rot) */
  float* out_xyz;                /* [n_pose][n_lig][3] minimised poses (mode 1) */
  double* out_energy;            /* [n_pose][4]: inter + intra, inter, intra (all with Vina's energy cap), affinity (kcal/mol) */
  double* out_terms;             /* [n_pose][5] unweighted intermolecular term sums of the INPUT pose (`## ligand` line of
                                    smina --score_only), or NULL */
  int32_t* out_stats;            /* [n_pose][2]: BFGS steps, energy evaluations, or NULL */
} B200Vina;
int b200dock_vina(B200Handle* h, const B200Vina* v, void* stream);

/* Introspection for tests / benchmarks: edge counts of the last evaluation
 * [E_ll, E_aa, E_al(=E_la), E_tor, E_sc], kernels launched by the last call, device time of the
 * dominant (tensor-product) kernel accumulated with CUDA events when profiling is enabled. */
int b200dock_last_edge_counts(B200Handle* h, int64_t counts[5]);
int b200dock_last_launch_count(const B200Handle* h, int64_t* n);
int b200dock_set_profiling(B200Handle* h, int enable);
int b200dock_tp_kernel_time_ms(B200Handle* h, double* ms, int64_t* launches);

/* Debug taps (tests only): copy an internal device buffer of the last evaluation to the host.
 * what: see B200_TAP_*; returns number of floats/ints written (<= cap) through *n. */
#define B200_TAP_H_LIG 0      /* [N_l][168] after the last conv layer */
#define B200_TAP_H_ATOM 1     /* [N_a][168] */
#define B200_TAP_EDGES 2      /* int32 pairs (s, d) of conv `arg` (0 lig, 1 atom, 2 al, 3 la, 4 tor, 5 sc) */
#define B200_TAP_H_LIG0 3     /* embeddings before layer 0 */
#define B200_TAP_H_ATOM0 4
#define B200_TAP_CONV_BUF 5   /* arg = conv*16 + which; which: 0 emb[E][48], 1 sh[E][9|8], 2 H1[E][160] (kernel 0),
                                 3 Zt[tiles][z][128] (kernel 0), 4 msg[E][168] (kernels 0, 5), 5 seg_ptr[T+2] (int32: first slot of
                                 every target, [T] = slots, [T+1] = real edges), 6 centre msg [N_l][12], 7 counts[T], 8 es[slots] */
/* Debug knobs (tests / tools only): key 0 = number of interaction layers to run (default 6); key 1 = wait-cycle accounting of the
 * fused conv kernels on (1) / off (0): 148 CTAs x 32 int64 counters (slots 0..7 MMA-issuing warp, 8..15 one fold warp; only filled by a
 * library built with -DB200DOCK_TRACE), read back with b200dock_debug_tap(what = 7); key 2 = node update in its warp-per-node
 * cross-check form (1) instead of the thread-per-(node, irreps block) kernel (0, default; bit-identical). */
int b200dock_debug_set(B200Handle* h, int key, int value);
int b200dock_debug_tap(B200Handle* h, int what, int arg, void* host_out, size_t cap_bytes, size_t* n_bytes);

#ifdef __cplusplus
}
#endif
#endif /* B200DOCK_H_ */
