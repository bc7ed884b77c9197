/* zpcb200 — C ABI of the B200-native (sm_100a) MPM transfer path and the parallel primitives it
 * rides on.  Drop-in boundary for zenustech/zpc's CUDA backend on this path.
 *
 * Conventions (mirroring the reference's own thin C layers):
 *   - plain pointers, sizes and POD views; no C++/torch types cross this boundary;
 *   - every entry returns int: 0 = ok, otherwise a cudaError_t value (or ZPCB200_E_* < 0 for misuse)
 *     — nothing throws (reference: u32 error code out of Cuda::launchKernel, cuda/Cuda.h:81-84,
 *     latched per context, cuda/Cuda.h:291-312);
 *   - every entry is asynchronous on the `stream` it is given (reference: pol.getStream(),
 *     cuda/execution/ExecutionPolicy.cuh:888-890); the library owns no streams, contexts or
 *     persistent memory;
 *   - scratch memory follows CUB's two-phase protocol that the reference already uses
 *     (ExecutionPolicy.cuh:803-812): call with temp == NULL to get *temp_bytes, then call again.
 *
 * All pointers are DEVICE pointers unless a name ends in _host.  Paths cited below are relative to
 * /root/reference/include/zensim/.
 */
#ifndef ZPCB200_H
#define ZPCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *zpc_stream_t; /* cudaStream_t */

#define ZPCB200_OK 0
#define ZPCB200_E_BADARG (-1)
#define ZPCB200_E_TEMP_TOO_SMALL (-2)
#define ZPCB200_E_UNSUPPORTED (-3)

/* ------------------------------------------------------------------------------------------ */
/* Iterator port — replaces py_interop/GenericIterator.hpp:11-16 (aosoa_iterator_port) 1:1.     */
/* Element k of the range lives at                                                              */
/*   base + ((( (idx+k) >> numTileBits) * numChns) << numTileBits | ((idx+k) & tileMask))       */
/* (GenericIterator.hpp:84-98).  A contiguous pointer is {ptr, 0, 0, 0, 1}.                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct zpc_port {
  void *base;
  uint32_t idx, numTileBits, tileMask, numChns;
} zpc_port;

/* ------------------------------------------------------------------------------------------ */
/* Parallel primitives — replace the CUB calls under CudaExecutionPolicy                        */
/* (cuda/execution/ExecutionPolicy.cuh: reduce :649-681, inclusive_scan :552-590,               */
/*  exclusive_scan :601-632, radix_sort_pair :755-826, radix_sort :828-866).                     */
/* <T> in {i32,u32,i64,f32,f64}; scans/reductions use the identities the reference's C ABI passes     */
/* (py_interop/cuda/ExecutionPolicy.cpp:41-90): 0 for sum, 1 for prod (:48-54), numeric max for min, lowest for max. */
/* `out` of a reduce is one element on the device.  Ranges are zpc_ports; n = last - first.      */
/* ------------------------------------------------------------------------------------------ */
#define ZPCB200_DECL_REDUCE_SCAN(S)                                                              \
  int zpcb200_reduce_sum_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, \
                             zpc_stream_t stream);                                               \
  int zpcb200_reduce_prod_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, \
                              zpc_stream_t stream);                                              \
  int zpcb200_reduce_min_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, \
                             zpc_stream_t stream);                                               \
  int zpcb200_reduce_max_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, \
                             zpc_stream_t stream);                                               \
  int zpcb200_exclusive_scan_sum_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out,  \
                                     size_t n, zpc_stream_t stream);                             \
  int zpcb200_inclusive_scan_sum_##S(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out,  \
                                     size_t n, zpc_stream_t stream);
ZPCB200_DECL_REDUCE_SCAN(i32)
ZPCB200_DECL_REDUCE_SCAN(u32)
ZPCB200_DECL_REDUCE_SCAN(i64)
ZPCB200_DECL_REDUCE_SCAN(f32)
ZPCB200_DECL_REDUCE_SCAN(f64)

/* Stable LSD radix sort on bits [sbit, ebit) of the key (ExecutionPolicy.hpp:765-781).  Signed
 * keys order as signed.  keys_in/vals_in are not modified; out may not alias in.  n <= 2^30. */
#define ZPCB200_DECL_SORT(S, KT)                                                                 \
  int zpcb200_radix_sort_pair_##S(void *temp, size_t *temp_bytes, zpc_port keys_in,              \
                                  zpc_port vals_in, zpc_port keys_out, zpc_port vals_out,        \
                                  size_t n, int sbit, int ebit, zpc_stream_t stream);            \
  int zpcb200_radix_sort_##S(void *temp, size_t *temp_bytes, zpc_port keys_in, zpc_port keys_out, \
                             size_t n, int sbit, int ebit, zpc_stream_t stream);
ZPCB200_DECL_SORT(u32, uint32_t)
ZPCB200_DECL_SORT(i32, int32_t)
ZPCB200_DECL_SORT(u64, uint64_t)

/* merge_sort / merge_sort_pair (cuda/execution/ExecutionPolicy.cuh:686-760; host twins execution/ExecutionPolicy.hpp:
 * 285-455): stable ascending sort under operator<, IN PLACE, <T> in {i32,f32,f64} like the reference's C layer
 * (py_interop/cuda/ExecutionPolicy.cpp:100-112) plus {u32,i64,u64} (the generic call takes any key type); values are i32.  -0.0 and +0.0 compare equal (their relative order is
 * kept); NaN keys sort by bit pattern. */
#define ZPCB200_DECL_MERGE_SORT(S)                                                                  \
  int zpcb200_merge_sort_pair_##S(void *temp, size_t *temp_bytes, zpc_port keys, zpc_port vals,     \
                                  size_t n, zpc_stream_t stream);                                   \
  int zpcb200_merge_sort_##S(void *temp, size_t *temp_bytes, zpc_port keys, size_t n, zpc_stream_t stream);
ZPCB200_DECL_MERGE_SORT(i32)
ZPCB200_DECL_MERGE_SORT(u32)
ZPCB200_DECL_MERGE_SORT(i64)
ZPCB200_DECL_MERGE_SORT(u64)
ZPCB200_DECL_MERGE_SORT(f32)
ZPCB200_DECL_MERGE_SORT(f64)

/* ------------------------------------------------------------------------------------------ */
/* View PODs (non-owning)                                                 */
/* ------------------------------------------------------------------------------------------ */
/* ParticlesView<cuda, Particles<f32,3>> — geometry/Structurefree.hpp:242-251,278-283.
 * AoS per attribute: X,V = vec3 (12 B stride), F,C = column-major vec9 (36 B), M,J,logJp scalars. */
typedef struct zpc_particles_view {
  float *M, *X, *V, *Dinv, *J, *F, *C, *logJp;
  size_t count;
} zpc_particles_view;

/* HashTableView<cuda, HashTable<i32,3,int>> — container/HashTable.hpp:332-347,487-490.
 * keys/activeKeys are packed vec3i (12 B); tableSize = next_2pow(expected) * 16 (:70,87-90). */
typedef struct zpc_hashtable_view {
  int *keys, *indices, *status; /* table_t */
  int *activeKeys;
  int tableSize;
  int *cnt;
} zpc_hashtable_view;

/* BHTView<cuda, bht<i32,3,int,16>> — container/Bht.hpp:459-482, 758-763: buckets of 16 slots, three universal hashes
 * (py_interop/HashUtils.hpp:7-47), keys stored as 16-byte padded vec3i slots (HashUtils.hpp:49-51) whose unused
 * value is the byte pattern 0x3f (Bht.hpp:108-112,126-131).  hf = {hf0.x, hf0.y, hf1.x, hf1.y, hf2.x, hf2.y}. */
typedef struct zpc_bht_view {
  int *keys; /* tableSize x 4 ints */
  int *indices, *status;
  int *activeKeys; /* packed vec3i */
  uint32_t tableSize, numBuckets;
  int *cnt, *success;
  uint32_t hf[6];
} zpc_bht_view;

/* SparseGridView<cuda, SparseGrid<3,f32,8>> — geometry/SparseGrid.hpp:239-243, 912-915: bht keyed by the block ORIGIN
 * in cell coordinates (:305-309) + TileVector<f32,512> (tile b = [numChannels][512] floats, cell offset
 * (x*8+y)*8+z, :275-283) + 4x4 index-to-world transform (row vector convention: world = (X,1) * M, :256-258) +
 * background value. */
typedef struct zpc_sparsegrid_view {
  zpc_bht_view table;
  float *grid;
  size_t numBlocks; /* capacity in blocks */
  int numChannels;
  float transform[16]; /* row-major M[i][j] */
  float background;
} zpc_sparsegrid_view;

/* The collocated grid of GridsView<cuda, Grids<f32,3,4>> — geometry/Structure.hpp:876-880: a
 * TileVector<f32,64> whose tile b holds block b as [numChannels][64] floats, channels
 * {m, v(3), rhs(3)} (simulation/mpm/Simulator.cpp:116), cell id (x<<4)|(y<<2)|z (:851-859). */
typedef struct zpc_grids_view {
  float *tiles;
  size_t numBlocks; /* capacity in blocks */
  int numChannels;  /* 7 */
  float dx;
} zpc_grids_view;

/* TileVectorUnnamedView<cuda, TileVector<f32,32>> — container/TileVector.hpp:693-723: AoSoA,
 * element (chn,i) at base[(i/32*numChannels + chn)*32 + i%32] (:108,768-769).  For particle bins the
 * channel order is fixed below (Structurefree.hpp:220 "mass, pos, vel, J, F, C, logJp" minus J/logJp). */
typedef struct zpc_tilevector_view {
  float *base;
  size_t size;
  int numChannels;
} zpc_tilevector_view;
enum { ZPC_PB_M = 0, ZPC_PB_X = 1, ZPC_PB_V = 4, ZPC_PB_C = 7, ZPC_PB_F = 16, ZPC_PB_NCH = 25 };

/* FixedCorotatedConfig — physics/ConstitutiveModel.hpp:725-742 */
typedef struct zpc_fixed_corotated {
  float rho, volume;
  int dim;
  float E, nu;
} zpc_fixed_corotated;

/* VonMisesFixedCorotatedConfig — physics/ConstitutiveModel.hpp:743-747 */
typedef struct zpc_vonmises_fixed_corotated {
  float rho, volume;
  int dim;
  float E, nu, yieldStress;
} zpc_vonmises_fixed_corotated;

/* DruckerPragerConfig — physics/ConstitutiveModel.hpp:748-757 (sand; the default yieldSurface is
 * sqrt(2/3) * 2 sin(30 deg) / (3 - sin(30 deg)), :756).  logJp0 and fa are carried like the reference's struct; the
 * transfer reads cohesion, beta, yieldSurface and volumeCorrection (P2G.hpp:94-96). */
typedef struct zpc_drucker_prager {
  float rho, volume;
  int dim;
  float E, nu, logJp0, fa, cohesion, beta;
  int volumeCorrection;
  float yieldSurface;
} zpc_drucker_prager;

/* NACCConfig — physics/ConstitutiveModel.hpp:758-776 (snow).  bulk() and Msqr() are derived from E, nu, fa and dim on
 * the host exactly as the struct's members do (fa goes to sin as is). */
typedef struct zpc_nacc {
  float rho, volume;
  int dim;
  float E, nu, logJp0, fa, xi, beta;
  int hardeningOn;
} zpc_nacc;

/* EquationOfStateConfig — physics/ConstitutiveModel.hpp:730-734 (gamma is forced to 7 by P2G.hpp:72-75) */
typedef struct zpc_equation_of_state {
  float rho, volume;
  int dim;
  float bulk, gamma, viscosity;
} zpc_equation_of_state;

/* model_kind of the entries that take any model through a pointer (zpcb200_g2p2g_apic, zpcb200_sg_p2g_apic_model) */
enum { ZPC_MODEL_FIXED_COROTATED = 0, ZPC_MODEL_VONMISES = 1, ZPC_MODEL_DRUCKER_PRAGER = 2, ZPC_MODEL_NACC = 3, ZPC_MODEL_EOS = 4 };

/* An analytic Collider — geometry/Collider.h:10-143 over AnalyticLevelSet<Plane> / <Sphere> / <Cuboid>
 * (geometry/AnalyticLevelSet.h:11-43, 130-157, 55-126) with its rigid motion x = R s X + b (Collider.h:16-24, 136-143):
 * translation b and its rate, rotation R (row-major) and angular velocity, uniform scale s and its rate.  The level set
 * is given in material space; the defaults (b = 0, R = I, s = 1, all rates 0) are a static collider.  Initialise with
 * zpcb200_collider_static() or set every field. */
enum { ZPC_GEOM_PLANE = 0, ZPC_GEOM_SPHERE = 1, ZPC_GEOM_CUBOID = 2 };
enum { ZPC_COLLIDER_STICKY = 0, ZPC_COLLIDER_SLIP = 1, ZPC_COLLIDER_SEPARATE = 2 }; /* collider_e, Collider.h:8 */
typedef struct zpc_collider {
  int geometry, type;
  float origin[3]; /* plane origin | sphere centre | cuboid min corner (material space) */
  float normal[3]; /* plane unit normal | {radius, -, -} | cuboid max corner */
  float b[3], dbdt[3];
  float R[9];
  float omega[3];
  float s, dsdt;
} zpc_collider;

/* ------------------------------------------------------------------------------------------ */
/* MPM path                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* partition_for_particles / CleanSparsity + ComputeSparsity + EnlargeSparsity{lo,hi}
 * (simulation/sparsity/SparsityOp.hpp:41-112, SparsityCompute.tpp:6-24); the reference's substep uses
 * enlarge {0,2} (every block the 3^3 stencil can touch).  {-1,3} adds one more ring so that the partition stays
 * sufficient while particles drift by up to a block between re-binnings.  x = port over vec3
 * positions (AoS: {X,0,0,0,3}; AoSoA bin channel: {base+ZPC_PB_X*32, 0, 5, 31, 25}).  Writes keys /
 * indices / status / activeKeys / *cnt so that the unmodified HashTableView::query (:447-456)
 * resolves every active block.  Block numbering is DETERMINISTIC here: index = rank of the block key
 * in lexicographic (x,y,z) order (the reference's is insertion-order and racy).
 * *overflow (device int, may be NULL) is set to 1 if the table is too small. */
int zpcb200_partition_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx,
                            zpc_hashtable_view table, int enlarge_lo, int enlarge_hi, int *overflow,
                            zpc_stream_t stream);
/* the same with 64-bit block codes: block coordinates in [-2^20, 2^20) per axis (cells beyond +-2 048), about twice the scratch;
 * the table, numbering rule (rank of the key) and every consumer are unchanged.  The default entry refuses such coordinates
 * through *overflow. */
int zpcb200_partition_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx,
                            zpc_hashtable_view table, int enlarge_lo, int enlarge_hi, int *overflow,
                            zpc_stream_t stream);

/* index_buckets_for_particles (simulation/particle/Query.tpp:9-58 = CleanSparsity + ComputeSparsity{blockLen 1, offset 0,
 * displacement} + SpatiallyCount + exclusive_scan + SpatiallyDistribute, SparsityOp.hpp:41-86, 115-195): `table` becomes the
 * table of occupied CELLS (floor(x/dx + displacement) per axis; the reference sizes it for n entries), counts / offsets get
 * n + 1 entries each (the reference allocates table.size() + 1; entries past that stay 0 / n), indices[offsets[b] ..
 * offsets[b+1]) are the particles of bucket b.  Deterministic where the reference is racy: bucket number = rank of the cell
 * key, ids inside a bucket ascending (what the reference's serial policy produces).  Cell coordinates in [-512, 511]
 * (flagged through *overflow otherwise).  n <= 2^30. */
int zpcb200_index_buckets_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, float displacement,
                                zpc_hashtable_view table, int *counts, int *offsets, int *indices, int *overflow,
                                zpc_stream_t stream);
int zpcb200_index_buckets_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, float displacement,
                                zpc_hashtable_view table, int *counts, int *offsets, int *indices, int *overflow,
                                zpc_stream_t stream);  /* 64-bit cell codes: cells in [-2^20, 2^20) instead of [-512, 511] */

/* CleanGridBlocks (simulation/grid/GridOp.hpp:54-69) over blocks [0, *cnt): every channel of every cell becomes 0.
 * ResetGrid (GridOp.hpp:166-182) is the same operation on one Grid of the Grids: this entry serves both. */
int zpcb200_clean_grid(zpc_grids_view grids, const int *cnt, zpc_stream_t stream);

/* P2GTransfer<apic, FixedCorotatedConfig> (simulation/transfer/P2G.hpp:45-127) on the reference's
 * own AoS Particles layout, any particle order.  One thread per particle, 27x7 float reds. */
int zpcb200_p2g_apic_fcr(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                         float dt, zpc_fixed_corotated model, zpc_stream_t stream);

/* ComputeGridBlockVelocity (GridOp.hpp:71-110): v = mv/m + extf*dt, *maxVelSqr = max |v|^2.
 * mode 0 = as shipped (rhs ignored); mode 1 = explicit update v = (mv + rhs)/m + extf*dt. */
int zpcb200_grid_update(zpc_grids_view grids, const int *cnt, float dt, const float extf_host[3],
                        int mode, float *maxVelSqr, zpc_stream_t stream);

/* GridMomentumToVelocity (simulation/grid/GridOp.hpp:184-214): for every cell of blocks [0, *cnt) whose mass (channel mChn) is not
 * zero, momentum channels mvChn..mvChn+2 become velocities (mv * (1/m)); *maxVelSqr = max(*maxVelSqr, |v|^2).  No gravity and no
 * rhs — that is zpcb200_grid_update (ComputeGridBlockVelocity). */
int zpcb200_grid_momentum_to_velocity(zpc_grids_view grids, const int *cnt, int mChn, int mvChn, float *maxVelSqr,
                                      zpc_stream_t stream);

/* GridAngularMomentum (GridOp.hpp:216-262): sum6[0..2] += sum over cells with mass of x cross mv, sum6[3..5] += sum of mv,
 * x = (blockkey*4 + cell coord) * dx; products in float, sums in double (device memory, six doubles, ADDED to: zero them first).
 * The reference adds six double atomics per cell in launch order; the sums here differ from it only by the order of the double
 * additions. */
int zpcb200_grid_angular_momentum(zpc_grids_view grids, zpc_hashtable_view table, int mChn, int mvChn, double *sum6,
                                  zpc_stream_t stream);

/* host helper: a collider with the default rigid motion */
zpc_collider zpcb200_collider_static(int geometry, int type, const float origin[3], const float normal_or_radius[3]);

/* ApplyBoundaryConditionOnGridBlocks (GridOp.hpp:112-164): for every cell with mass > 0 of blocks [0, *cnt), project
 * the grid velocity (channels 1-3) against the collider; node position = (blockkey*4 + cell coord) * dx. */
int zpcb200_apply_boundary(zpc_grids_view grids, zpc_hashtable_view table, zpc_collider collider,
                           zpc_stream_t stream);

/* P2GTransfer<apic, VonMisesFixedCorotatedConfig> (P2G.hpp:89-90 -> compute_stress_vonmisesfixedcorotated,
 * physics/ConstitutiveModel_Vol_dP.hpp:49-110): elastoplastic solid with a von Mises yield surface; G2P is
 * zpcb200_g2p_apic (the projection of F stays local to P2G, like in the reference). */
int zpcb200_p2g_apic_vonmises(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                              float dt, zpc_vonmises_fixed_corotated model, zpc_stream_t stream);

/* P2GTransfer<apic, DruckerPragerConfig> / <apic, NACCConfig> (P2G.hpp:92-102 -> compute_stress_sand /
 * compute_stress_nacc, physics/ConstitutiveModel_Vol_dP.hpp:116-326): plastic models with a per-particle logJp
 * (pars.logJp must be set) that the return mapping reads and P2G writes back; F is not modified (the projection stays
 * local, like in the reference); G2P is zpcb200_g2p_apic. */
int zpcb200_p2g_apic_drucker_prager(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                                    float dt, zpc_drucker_prager model, zpc_stream_t stream);
int zpcb200_p2g_apic_nacc(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                          float dt, zpc_nacc model, zpc_stream_t stream);

/* ComputeGridBlockVelocity + ApplyBoundaryConditionOnGridBlocks for up to ZPCB200_MAX_COLLIDERS static colliders in one
 * pass over the grid (same results as zpcb200_grid_update followed by zpcb200_apply_boundary per collider, in order;
 * *maxVelSqr is taken before the projection, like the reference's sequence).  colliders_host: host array. */
#define ZPCB200_MAX_COLLIDERS 4
int zpcb200_grid_update_bc(zpc_grids_view grids, zpc_hashtable_view table, float dt, const float extf_host[3],
                           int mode, const zpc_collider *colliders_host, int ncolliders, float *maxVelSqr,
                           zpc_stream_t stream);

/* G2P2GTransfer<apic, Model> (simulation/transfer/G2P2G.hpp:49-141) — the matrix-free force evaluation of the implicit solver
 * (simulation/mpm/ImplicitMPM.hpp:32-59): gridv / gridr are the DOF vectors of dof_view<space, 3>(Vector<float>), three floats
 * per node, node = blockno * 64 + cellid.  C is gathered from gridv, the trial F = (I + dt C) F stays in registers (particles are
 * not modified, logJp is read only), W * (P F^T vol * D_inv) * (x_i - x_p) is ADDED to gridr (clear it first, like DofFill).
 * model_kind selects the struct `model` points to (host memory).  Parity: against the reference's own functor run on cuda_exec() with
 * a plain three-floats-per-node DOF view (oracle/ref_driver_cuda.cu; its dof_view types do not compile under gcc 13, the functor only
 * needs get / ref) for the fixed-corotated and von Mises models, against the restated oracle zo_g2p2g for all five. */
int zpcb200_g2p2g_apic(zpc_particles_view pars, zpc_hashtable_view table, float dx, float dt, int model_kind, const void *model,
                       const float *gridv, float *gridr, zpc_stream_t stream);

/* G2PTransfer<apic> (simulation/transfer/G2P.hpp:43-84), AoS layout, any order. */
int zpcb200_g2p_apic(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                     float dt, zpc_stream_t stream);

/* The same two functors for EquationOfStateConfig (weakly compressible fluid: per-particle J instead of F;
 * P2G.hpp:66-87, G2P.hpp:69-73).  pars.J must be set; F is not touched. */
int zpcb200_p2g_apic_eos(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                         float dt, zpc_equation_of_state model, zpc_stream_t stream);
int zpcb200_g2p_apic_eos(zpc_particles_view pars, zpc_hashtable_view table, zpc_grids_view grids,
                         float dt, zpc_stream_t stream);

/* ---- LBvh<3, int, f32> on the primitives -------------------------------------------------- */
/* LBvhView / LBvh members (container/Bvh.hpp:173-174, 497-510): boxes are AABBBox<3,f32> = {min[3], max[3]} (six floats);
 * orderedBvs / auxIndices / parents / levels hold 2n-1 nodes in DFS pre-order (n nodes when n <= 2), leafInds n entries.
 * auxIndices = primitive id at a leaf, escape index at an internal node; levels = 0 at a leaf, else the length of the chain
 * of left children below the node. */
typedef struct zpc_lbvh_view {
  float *orderedBvs;
  int *auxIndices, *parents, *levels, *leafInds;
} zpc_lbvh_view;
/* LBvh::build(policy, primBvs, wrapv<Refit>) (Bvh.hpp:835-1000): whole box (reduce) -> Morton codes -> radix_sort_pair ->
 * Karras topology -> exclusive_scan -> DFS layout [-> refit].  Two-phase temp.  numLeaves <= 2^30.  Arrays are
 * bit-identical to the reference's (the topology is a function of the sorted codes). */
int zpcb200_lbvh_build(void *temp, size_t *temp_bytes, const float *primBvs, size_t numLeaves, zpc_lbvh_view bvh,
                       int refit, zpc_stream_t stream);
/* Batched LBvhView::iter_neighbors (Bvh.hpp:660-689, stack-free traversal with escape indices): one query box per thread.
 * out == NULL: counts[q] = number of primitives whose box overlaps query q.  Otherwise their ids are written to
 * out[offsets[q] ...] in visiting order — the usual count -> zpcb200_exclusive_scan_sum_i32 -> fill sequence. */
int zpcb200_lbvh_query(zpc_lbvh_view bvh, size_t numLeaves, const float *queryBvs, size_t numQueries, int *counts,
                       const int *offsets, int *out, zpc_stream_t stream);
/* LBvh::refit (Bvh.hpp:1229-1259): new boxes for the same primitives, same topology. */
int zpcb200_lbvh_refit(void *temp, size_t *temp_bytes, const float *primBvs, size_t numLeaves, zpc_lbvh_view bvh,
                       zpc_stream_t stream);

/* ---- SparseGrid<3,f32,8> variant of the path ----------------------------------------------- */
/* Host helpers mirroring the bht constructor: the three universal hashes it draws from std::mt19937(2)
 * (Bht.hpp:165-169, Bcht.hpp:39-43) and evaluateTableSize (Bht.hpp:154-158). */
void zpcb200_bht_params(uint32_t hf_host[6]);
size_t zpcb200_bht_table_size(size_t expected_entries);
/* The MPM functors need an axis-aligned uniform transform without translation (world = X * dx): dx = transform[0].
 * Otherwise ZPCB200_E_UNSUPPORTED. */
/* Partition for particles on side-8 blocks: the ComputeSparsity / EnlargeSparsity convention of
 * simulation/sparsity/SparsityOp.hpp:58-112 with blockLen 8; table keys = block origins.  Writes keys / indices /
 * activeKeys / *cnt / *success so that the unmodified BHTView::query (Bht.hpp:666-700) resolves every active block;
 * a key sits in the first of its three candidate buckets that has room (<= 15 keys, threshold 14).  Block numbering
 * is deterministic: index = rank of the key in lexicographic order. */
int zpcb200_sg_partition_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, zpc_sparsegrid_view sg,
                               int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream);
int zpcb200_sg_partition_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, zpc_sparsegrid_view sg,
                               int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream);  /* 64-bit block codes, as above */
/* CleanGridBlocks / P2GTransfer<apic, FixedCorotated> / ComputeGridBlockVelocity / G2PTransfer<apic> on the SparseGrid
 * (same arithmetic as the Grids<f32,3,4> entries above; channels {m, v(3), rhs(3)}), AoS particles in any order. */
int zpcb200_sg_clean(zpc_sparsegrid_view sg, zpc_stream_t stream);
int zpcb200_sg_p2g_apic_fcr(zpc_particles_view pars, zpc_sparsegrid_view sg, float dt, zpc_fixed_corotated model,
                            zpc_stream_t stream);
/* P2GTransfer with any of the five constitutive models on the SparseGrid (model_kind = ZPC_MODEL_*, model -> the matching struct,
 * host memory; J / logJp in the particle view where the model needs them) and the J-variant of G2PTransfer. */
int zpcb200_sg_p2g_apic_model(zpc_particles_view pars, zpc_sparsegrid_view sg, float dt, int model_kind, const void *model,
                              zpc_stream_t stream);
int zpcb200_sg_g2p_apic_eos(zpc_particles_view pars, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream);
int zpcb200_sg_grid_update(zpc_sparsegrid_view sg, float dt, const float extf_host[3], int mode, float *maxVelSqr,
                           zpc_stream_t stream);
int zpcb200_sg_g2p_apic(zpc_particles_view pars, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream);
/* SparseGridView::valueOr(false_c, chn, indexCoord, default) (SparseGrid.hpp:340-351) at n integer coordinates
 * (coords = n x 3 ints on the device); absent blocks give dflt. */
int zpcb200_sg_value_or(zpc_sparsegrid_view sg, int chn, const int *coords, size_t n, float dflt, float *out,
                        zpc_stream_t stream);
/* iCoord(bno, cno) and wCoord(bno, cno) (SparseGrid.hpp:407-416) for n (block, cell) pairs; either output may be NULL. */
int zpcb200_sg_cell_coords(zpc_sparsegrid_view sg, const int *bno, const int *cno, size_t n, int *icoord,
                           float *wcoord, zpc_stream_t stream);

/* Renumbering utilities.
 * TileVector::reorderTiles (container/TileVector.hpp:641-691): scatter != 0: dst tile map[i] <- src tile i; gather:
 * dst tile i <- src tile map[i].  tileLength = 32 | 64 | 512 ...; src != dst (the reference also builds a new vector). */
int zpcb200_tilevector_reorder_tiles(const float *src, float *dst, int numChannels, int tileLength,
                                     const int *map, size_t numTiles, int scatter, zpc_stream_t stream);
/* bht::reorder (container/Bht.hpp:343-400): writes the reordered key list to orderedKeys (n x vec3i; the caller makes it
 * the table's activeKeys, as the reference's move-assignment does) and renumbers the index stored with every key. */
int zpcb200_bht_reorder(zpc_bht_view table, const int *map, int n, int scatter, int *orderedKeys,
                        zpc_stream_t stream);
/* Gather map that puts the n active blocks of a SparseGrid in Morton (Z-curve) order of their block coordinates:
 * map[i] = current index of the block that becomes block i — 30-bit codes, radix_sort_pair of (code, index).  Feed it
 * to zpcb200_bht_reorder and zpcb200_tilevector_reorder_tiles (gather).  *overflow is set if a block lies outside
 * [-512, 511] blocks per axis. */
int zpcb200_sg_morton_order(void *temp, size_t *temp_bytes, zpc_sparsegrid_view sg, int n, int *map,
                            int *overflow, zpc_stream_t stream);

/* ---- binned (block-sorted AoSoA) fast path ------------------------------------------------- */
/* Bins: particles sorted by home block (the block ComputeSparsity assigns, SparsityOp.hpp:68-79),
 * stored densely in a 25-channel TileVector<f32,32>; bin b covers particles
 * [binStart[b], binStart[b+1]) and has home block coordinates binKey[3b..3b+2].  A block with more
 * than ZPCB200_BIN_MAX particles is split into several bins. */
#define ZPCB200_BIN_MAX 1024
#define ZPCB200_CELL_GROUPS 217     /* 6x6 columns x 6 z-levels of home cells around a block + 1 far-stray group */
#define ZPCB200_CELL_GROUPS_PAD 224 /* row stride of cellStart */
typedef struct zpc_bins_view {
  zpc_tilevector_view pars; /* numChannels == ZPC_PB_NCH */
  int *binStart;            /* [binCapacity + 1] */
  int *binKey;              /* [binCapacity * 3] */
  int *numBins;             /* device scalar */
  int binCapacity;
  /* Optional cell-order cache (all three NULL to disable).  The binned G2P knows every particle's NEW home cell,
   * so it leaves, per bin, the particle slots grouped by (column, z) for the next binned P2G, which then skips its
   * own in-kernel counting sort.  cellOrderValid is set (non-zero) by the binned G2P and cleared by
   * bin_particles / rebin_particles; anything else that moves particles must clear it. */
  unsigned short *cellOrder; /* [pars.size]: k-th particle of its bin in group order -> slot relative to binStart */
  unsigned short *cellStart; /* [binCapacity * ZPCB200_CELL_GROUPS_PAD]: group offsets, entry 217 = bin count */
  int *cellOrderValid;       /* device flag */
  /* Optional status word (device int, may be NULL; the caller zeroes it): every capacity / consistency condition of the binned
   * path ORs a ZPC_BINS_* bit into it instead of failing silently.  Nothing is thrown; read it where the host synchronises
   * anyway (MpmSolver does at every re-bin). */
  int *status;
} zpc_bins_view;
enum {
  ZPC_BINS_HOME_BLOCK_MISSING = 1,    /* bin / rebin: a particle's home block is not in the table (filed under block 0) */
  ZPC_BINS_BIN_CAPACITY = 2,          /* bin / rebin: more bins than binCapacity — numBins is set to 0, nothing will be transferred */
  ZPC_BINS_BLOCK_CAPACITY = 4,        /* bin / rebin: the table holds more blocks than binCapacity (blocks beyond it are dropped) */
  ZPC_BINS_STENCIL_BLOCK_MISSING = 8, /* binned P2G / G2P: a block a particle's stencil reaches is absent from the partition: that
                                         part of its mass / momentum is not transferred.  With partition = "with_rebin" this means a
                                         particle drifted by more than the extra ring since the last re-bin (re-bin more often) */
  ZPC_BINS_TMA_TIMEOUT = 16           /* binned G2P: a TMA bulk copy did not complete within 2^20 timed-out waits (cannot happen with
                                         consistent views); the kernel went on instead of trapping or hanging — results are invalid */
};

/* Sort AoS particles into bins (radix_sort_pair on the block rank + gather into AoSoA).  Requires a
 * partition built from the same positions.  order_out (may be NULL) receives the permutation:
 * binned particle i came from AoS particle order_out[i]. */
int zpcb200_bin_particles(void *temp, size_t *temp_bytes, zpc_particles_view pars,
                          zpc_hashtable_view table, float dx, zpc_bins_view bins, int *order_out,
                          zpc_stream_t stream);
/* Re-bin an already binned set in place after particles moved (src -> dst, both AoSoA). */
int zpcb200_rebin_particles(void *temp, size_t *temp_bytes, zpc_bins_view src,
                            zpc_hashtable_view table, float dx, zpc_bins_view dst,
                            zpc_stream_t stream);
/* dst[i] = src[idx[i]], i < n: permutes a per-particle side array (logJp, J) with the order a re-bin returned. */
int zpcb200_gather_f32(const float *src, const int *idx, float *dst, size_t n, zpc_stream_t stream);
/* Copy binned AoSoA particles back to the AoS view, slot i -> pars[i]. */
int zpcb200_unbin_particles(zpc_bins_view bins, zpc_particles_view pars, zpc_stream_t stream);

/* Same functors as above on the binned layout: one CTA per bin, the 2x2x2 block arena staged in
 * shared memory, cell-grouped register accumulation, bulk reduce-add of whole grid tiles. */
int zpcb200_p2g_apic_fcr_binned(zpc_bins_view bins, zpc_hashtable_view table, zpc_grids_view grids,
                                float dt, zpc_fixed_corotated model, zpc_stream_t stream);
/* P2GTransfer<apic, VonMisesFixedCorotatedConfig> on the binned layout: the same kernel, the model enters the record
 * phase only (P2G.hpp:89-90); G2P is zpcb200_g2p_apic_binned. */
int zpcb200_p2g_apic_vonmises_binned(zpc_bins_view bins, zpc_hashtable_view table, zpc_grids_view grids,
                                     float dt, zpc_vonmises_fixed_corotated model, zpc_stream_t stream);
/* DruckerPragerConfig / NACCConfig on the binned layout: logJp = one float per particle in BIN order (slot i of the bins), read
 * and written back by the record phase (P2G.hpp:93,101).  The caller permutes it together with the particles:
 * zpcb200_bin_particles(order_out) / zpcb200_rebin_particles_ordered(order_out) return the permutation (dst slot i <- src
 * order_out[i]). */
int zpcb200_p2g_apic_drucker_prager_binned(zpc_bins_view bins, float *logJp, zpc_hashtable_view table, zpc_grids_view grids,
                                           float dt, zpc_drucker_prager model, zpc_stream_t stream);
int zpcb200_p2g_apic_nacc_binned(zpc_bins_view bins, float *logJp, zpc_hashtable_view table, zpc_grids_view grids, float dt,
                                 zpc_nacc model, zpc_stream_t stream);
/* EquationOfStateConfig on the binned layout (P2G.hpp:66-87, G2P.hpp:69-73): J = one float per particle in bin order; the F
 * channels of the bins are neither read nor written. */
int zpcb200_p2g_apic_eos_binned(zpc_bins_view bins, const float *J, zpc_hashtable_view table, zpc_grids_view grids, float dt,
                                zpc_equation_of_state model, zpc_stream_t stream);
int zpcb200_g2p_apic_eos_binned(zpc_bins_view bins, float *J, zpc_hashtable_view table, zpc_grids_view grids, float dt,
                                zpc_stream_t stream);
int zpcb200_rebin_particles_ordered(void *temp, size_t *temp_bytes, zpc_bins_view src, zpc_hashtable_view table, float dx,
                                    zpc_bins_view dst, int *order_out, zpc_stream_t stream);
int zpcb200_g2p_apic_binned(zpc_bins_view bins, zpc_hashtable_view table, zpc_grids_view grids,
                            float dt, zpc_stream_t stream);

/* Kernel variants of the two binned functors (same results up to fp32 re-association; kept selectable so that each
 * can be measured and parity-tested).  p2g_sweep: 4 = a warp sweeps three cells at a time, lanes = 3 cells x 9 (x,y)
 * node columns, three z-nodes per lane; the fixed-corotated stress skips the Jacobi sweeps a whole warp has converged on (default);
 * 5 = the same sweep on packed fp32 arithmetic (FFMA2); 6 = the atomic-free plane sweep (lanes = 10 cells x 3 x-planes, a private
 * arena copy per warp region, no shared-memory atomics); 3 = one cell at a time, lanes = the 27 nodes, the reference's four Jacobi
 * sweeps always.  g2p_staged: 1 = the
 * particle channels G2P reads are staged with TMA bulk copies in 64-thread CTAs (default; 128 / 256 select that CTA
 * size instead); 0 = plain loads, 256-thread CTAs.
 * -1 leaves a setting unchanged.  Environment defaults: ZPCB200_P2G_SWEEP, ZPCB200_G2P_STAGED.  Not thread-safe
 * against concurrent launches. */
int zpcb200_set_tuning(int p2g_sweep, int g2p_staged);
int zpcb200_get_tuning(int *p2g_sweep_host, int *g2p_staged_host);

/* Block-binned fast path on the SparseGrid (round 2).  The bins of zpcb200_bin_particles hold the particles of one 4^3-cell region;
 * on side-8 blocks that region is one OCTANT of a block, so the same bins, cell-order cache and kernels serve both grids: the
 * arena's eight [7][64] tiles are the octants {b, b+1}^3, read with 128-bit loads and added back with 128-bit vector reductions
 * (an octant row of four z-cells is 16 contiguous bytes of the [numChannels][512] tile; Grids<f32,3,4> tiles are contiguous and
 * go through TMA).  bins.binCapacity bounds 8 x the number of active blocks.  sg.numChannels >= 7 (P2G) / >= 4 (G2P); channel
 * layout {m, v(3), rhs(3)} like the entries above.  order_out as in zpcb200_bin_particles / _rebin_particles_ordered. */
int zpcb200_sg_bin_particles(void *temp, size_t *temp_bytes, zpc_particles_view pars, zpc_sparsegrid_view sg, zpc_bins_view bins,
                             int *order_out, zpc_stream_t stream);
int zpcb200_sg_rebin_particles(void *temp, size_t *temp_bytes, zpc_bins_view src, zpc_sparsegrid_view sg, zpc_bins_view dst,
                               int *order_out, zpc_stream_t stream);
int zpcb200_sg_p2g_apic_fcr_binned(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, zpc_fixed_corotated model,
                                   zpc_stream_t stream);
int zpcb200_sg_g2p_apic_binned(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream);
/* the other constitutive models on the binned SparseGrid path: model_kind = ZPC_MODEL_*, model -> the matching struct (host memory);
 * scalar = logJp (Drucker-Prager, NACC: read and written back) or J (equation of state: read) per particle in BIN order — permute it
 * with the order a re-bin returns (zpcb200_gather_f32) —, NULL for the F-only models.  zpcb200_sg_g2p_apic_eos_binned advances J. */
int zpcb200_sg_p2g_apic_model_binned(zpc_bins_view bins, float *scalar, zpc_sparsegrid_view sg, float dt, int model_kind,
                                     const void *model, zpc_stream_t stream);
int zpcb200_sg_g2p_apic_eos_binned(zpc_bins_view bins, float *J, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream);

/* ---- multi-GPU one-ring halo, fused (SURVEY §8(e), §2.1 last paragraph; no reference counterpart) --------------------------
 * Every rank owns a receive buffer that its peers can address (peer-mapped over NVLink: torch symmetric memory, cudaIpc or
 * cuMem exports — the library only sees addresses): [2 halves][world senders][seg tiles][7 x 64 floats], all zero outside the
 * window between a send and its receive.
 *   send    = inside the binned P2G's write-back: a CTA that adds an arena tile to a grid block which another rank also holds
 *             issues one more TMA bulk reduce-add (cp.reduce.async.bulk ... add.f32, 1 792 B) for it — into the slot of that
 *             block in the PEER's buffer.  No pack kernel, no staging copy; the transfer overlaps the rest of the P2G.
 *   barrier = the caller's (one device-side barrier between the P2G and the grid update: symmetric-memory barrier, NCCL, ...).
 *   receive = inside the grid update: a shared block adds its slots (ascending peer rank: a fixed summation order), writes the
 *             complete m / rhs back like a single-GPU grid would hold them, zeroes the slots, then computes v.
 * half alternates every substep, so one barrier per substep suffices.
 * The maps come from zpcb200_halo_codes (this rank's block keys as sortable codes, padded) -> the caller's all_gather ->
 * zpcb200_halo_build: both ranks of a pair enumerate their common blocks in ascending key order, so the slot of a block is its
 * rank in that intersection — the same number on both sides, no negotiation, nothing read back to the host. */
#define ZPCB200_HALO_MAX_PEERS 16
#define ZPCB200_HALO_K 4 /* a block is shared with at most this many other ranks (slab or block-range shards: 1..3) */
typedef struct zpc_halo_view {
  const int *peer; /* [numBlocks * ZPCB200_HALO_K]: ranks that also hold block b, ascending, -1 terminated; NULL = no halo */
  const int *pos;  /* [numBlocks * ZPCB200_HALO_K]: slot of block b in the segment this rank and that peer share */
  int world, rank;
  int seg;  /* tiles per (half, sender) segment */
  int half; /* 0 | 1 */
  float *recv;                          /* this rank's receive buffer */
  float *peers[ZPCB200_HALO_MAX_PEERS]; /* every rank's receive buffer as addressable from this rank (peers[rank] == recv) */
  int *status;                          /* device int, ORed with ZPC_HALO_* (may be NULL) */
} zpc_halo_view;
enum { ZPC_HALO_TOO_MANY_PEERS = 1 /* a block is shared with more than ZPCB200_HALO_K ranks */, ZPC_HALO_SEGMENT_FULL = 2 /* > seg common blocks */ };
/* codes[i] = sortable 63-bit code of active block i (i < *cnt), INT64_MAX beyond: what the caller all_gathers ([capacity] each) */
int zpcb200_halo_codes(zpc_hashtable_view table, int capacity, long long *codes, zpc_stream_t stream);
/* all_codes: [world][capacity] gathered codes (each row ascending).  Writes peer / pos ([capacity * ZPCB200_HALO_K] ints each). */
int zpcb200_halo_build(void *temp, size_t *temp_bytes, const long long *all_codes, int world, int rank, int capacity, int seg,
                       int *peer_out, int *pos_out, int *status, zpc_stream_t stream);
/* P2GTransfer on the binned layout with the halo send fused into the write-back (fixed-corotated) */
int zpcb200_p2g_apic_fcr_binned_halo(zpc_bins_view bins, zpc_hashtable_view table, zpc_grids_view grids, float dt,
                                     zpc_fixed_corotated model, zpc_halo_view halo, zpc_stream_t stream);
/* ComputeGridBlockVelocity (+ up to ZPCB200_MAX_COLLIDERS colliders) with the halo receive fused in */
int zpcb200_grid_update_halo(zpc_grids_view grids, zpc_hashtable_view table, float dt, const float extf[3], int mode,
                             const zpc_collider *colliders, int ncolliders, float *maxVelSqr, zpc_halo_view halo, zpc_stream_t stream);

/* ---- multi-GPU one-ring halo, unfused building blocks (pack / exchange / unpack; the NCCL transport uses them) ------------- */
/* pack: copy tiles listed in blockIds[0..n) into a contiguous buffer (nch channels from chn0);
 * unpack_add / unpack_set: add / overwrite them back.  The exchange itself is NCCL send/recv (host). */
int zpcb200_halo_pack(zpc_grids_view grids, const int *blockIds, int n, int chn0, int nch,
                      float *buf, zpc_stream_t stream);
int zpcb200_halo_unpack_add(zpc_grids_view grids, const int *blockIds, int n, int chn0, int nch,
                            const float *buf, zpc_stream_t stream);
int zpcb200_halo_unpack_set(zpc_grids_view grids, const int *blockIds, int n, int chn0, int nch,
                            const float *buf, zpc_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Reference-style object ABI (py_interop/cuda/ExecutionPolicy.cpp:8-9, 39-134): an opaque policy   */
/* carrying {device, stream, sync} and primitives named <op>__b200_<T>_1 taking iterator ports.   */
/* Scratch comes from the stream-ordered pool (cudaMallocAsync), like ExecutionPolicy.cuh:806-815. */
/* ------------------------------------------------------------------------------------------ */
typedef struct zpcb200_policy zpcb200_policy;
zpcb200_policy *policy__b200(void);
void del_policy__b200(zpcb200_policy *);
void policy_set__b200(zpcb200_policy *, int device, zpc_stream_t stream, int sync);
int policy_last_error__b200(const zpcb200_policy *);
#define ZPCB200_DECL_POLICY_PRIMS(T, CT)                                                          \
  void reduce_sum__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last, zpc_port out);   \
  void reduce_prod__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last, zpc_port out);  \
  void reduce_min__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last, zpc_port out);   \
  void reduce_max__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last, zpc_port out);   \
  void exclusive_scan_sum__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last,          \
                                        zpc_port out);                                            \
  void inclusive_scan_sum__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last,          \
                                        zpc_port out);
ZPCB200_DECL_POLICY_PRIMS(int, int32_t)
ZPCB200_DECL_POLICY_PRIMS(float, float)
ZPCB200_DECL_POLICY_PRIMS(double, double)
#define ZPCB200_DECL_POLICY_MERGE(T)                                                              \
  void merge_sort__b200_##T##_1(zpcb200_policy *, zpc_port first, zpc_port last);                 \
  void merge_sort_pair__b200_##T##_1(zpcb200_policy *, zpc_port keys, zpc_port vals, size_t count);
ZPCB200_DECL_POLICY_MERGE(int)
ZPCB200_DECL_POLICY_MERGE(float)
ZPCB200_DECL_POLICY_MERGE(double)
void radix_sort__b200_int_1(zpcb200_policy *, zpc_port first, zpc_port last, zpc_port out);
void radix_sort_pair__b200_int_1(zpcb200_policy *, zpc_port keysIn, zpc_port valsIn,
                                 zpc_port keysOut, zpc_port valsOut, size_t count);

/* library info */
const char *zpcb200_version(void);
int zpcb200_kernel_launch_count(void); /* launches issued by this library since load (bench.py) */

#ifdef __cplusplus
}
#endif
#endif
