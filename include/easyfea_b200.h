/*
 * easyfea_b200 — C ABI of the B200-native (sm_100a, FP64) element-integration + CSR-assembly path of EasyFEA.
 *
 * The reference (matnoel/EasyFEA v3.5.1) is pure Python: it has no FFI layer for this path.  The entry points
 * below are what a ctypes binding inside the reference would call in place of its NumPy expressions; each one
 * cites the reference function it replaces (paths relative to the reference root).  See INTEGRATION.md for the
 * reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`; all floating point is FP64;
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); calls are asynchronous;
 *  - return value: 0 on success, non-zero on error (message via efb_last_error());
 *  - arrays are C-contiguous with the shapes written beside them; element axis `Ne`, Gauss axis `nPg`,
 *    Kelvin-Mandel strain size ns = 3 (dim 2) or 6 (dim 3), ndof = nPe*dof_n.
 *  - no torch / numpy types cross this boundary.
 */
#ifndef EASYFEA_B200_H
#define EASYFEA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------ */
const char* efb_last_error(void);
int efb_version(void);
/* number of CUDA devices visible, or -1 when the runtime cannot initialise (no GPU): never throws */
int efb_device_count(void);

/* One element group on the device: EasyFEA/FEM/_group_elem.py:41 (`_GroupElem`) reduced to what the path reads.
 * Tables come from `Get_dN_pg/Get_N_pg/Get_weight_pg(matrixType)` (:1063, :990, :788). */
typedef struct efb_group {
    int32_t dim;           /* 2 or 3 */
    int32_t nPe;           /* nodes per element */
    int32_t nPg;           /* Gauss points of the matrixType in use */
    int32_t coord_stride;  /* doubles per row of `coord` (3 in the reference) */
    int64_t Ne;
    const int32_t* connect; /* (Ne, nPe) rows of `coord` */
    const double* coord;    /* (Nn, coord_stride) */
    const double* dN_pg;    /* (nPg, dim, nPe) */
    const double* N_pg;     /* (nPg, nPe) */
    const double* w_pg;     /* (nPg) */
} efb_group;

/* broadcast modes of a coefficient, FeArray.broadcast EasyFEA/FEM/_linalg.py:426-476 */
enum { EFB_COEF_SCALAR = 0, EFB_COEF_E = 1, EFB_COEF_PG = 2, EFB_COEF_E_PG = 3 };
/* leading axes of a tensor coefficient (C, A): (), (Ne,), (Ne,nPg) */
enum { EFB_TENSOR_CONST = 0, EFB_TENSOR_E = 1, EFB_TENSOR_E_PG = 2 };

/* ---- G2-G8: Gauss-point geometry, _group_elem.py:832-1312 -------------------------------------------- */
/* Any output pointer may be NULL.  F (Ne,nPg,dim,dim) Get_F_e_pg :832; detF (Ne,nPg) signed, jac = |detF|
 * Get_jacobian_e_pg :871; wJ Get_weightedJacobian_e_pg :890; invF Get_invF_e_pg :902;
 * dN (Ne,nPg,dim,nPe) Get_dN_e_pg :1083; B (Ne,nPg,ns,nPe*dim) Get_B_e_pg :1241. */
int efb_geometry(const efb_group* g, double* F, double* detF, double* jac, double* wJ, double* invF, double* dN,
                 double* B, void* stream);

/* G8-G10, the cached per-Gauss-point factors (_group_elem.py:1314-1407); any output may be NULL:
 * leftDisp (Ne,nPg,nPe*dim,ns) = wJ B^T  Get_leftDispPart_e_pg :1315;  reaction (Ne,nPg,nPe*dof_n,nPe*dof_n) = wJ N^T N
 * with the block-diagonal N of Get_N_pg_rep  Get_ReactionPart_e_pg :1338;  diffuse (Ne,nPg,nPe,dim) = wJ dN^T
 * Get_DiffusePart_e_pg :1363;  source (Ne,nPg,nPe*dof_n,dof_n) = wJ N^T  Get_SourcePart_e_pg :1383.  dof_n in 1..3. */
int efb_geometry_parts(const efb_group* g, int dof_n, double* leftDisp, double* reaction, double* diffuse, double* source,
                       void* stream);

/* ---- O1-O4: operators, EasyFEA/FEM/Operators/Bilinear.py, Linear.py ---------------------------------- */
/* LinearizedElasticity Bilinear.py:62-79: out (Ne,ndof,ndof) = scale * sum_p wJ B^T C B; C (ns,ns) with leading
 * axes per C_mode (device).  A homogeneous C (EFB_TENSOR_CONST) is passed to the kernel by value: give its ns*ns
 * values in C_host (HOST pointer; then C may be NULL), otherwise they are fetched from C with a stream sync. */
int efb_elastic_Ke(const efb_group* g, const double* C, const double* C_host, int C_mode, double scale, double* out,
                   void* stream);
/* UV Bilinear.py:42-59: out (Ne,ndof,ndof) = scale * sum_p coef wJ N^T N, N block-diagonal for dof_n>1.
 * coef: device array per coef_mode, or NULL with the value in coef_scalar. */
int efb_mass_Me(const efb_group* g, const double* coef, int coef_mode, double coef_scalar, int dof_n, double scale,
                double* out, void* stream);
/* GradUGradV :25-39 (A == NULL) and GradU_A_GradV :229-249: out (Ne,nPe,nPe) = scale * sum_p coef wJ dN^T A dN. */
int efb_diffusion_Ke(const efb_group* g, const double* A, int A_mode, const double* coef, int coef_mode,
                     double coef_scalar, double scale, double* out, void* stream);
/* Linear.V Linear.py:18-35: out (Ne, nPe*dof_n, dof_n) = scale * sum_p f wJ N^T (the reference keeps the last axis). */
int efb_source_Fe(const efb_group* g, const double* f, int f_mode, double f_scalar, int dof_n, double scale,
                  double* out, void* stream);
/* Linear.InternalForce Linear.py:38-52: out (Ne,ndof) = sum_p wJ B^T sigma, sigma (Ne,nPg,ns). */
int efb_internal_force(const efb_group* g, const double* sigma, double* out, void* stream);
/* Calc_Epsilon_e_pg EasyFEA/Models/Elastic/_laws.py:127-157 with Locates_sol_e _group_elem.py:1769:
 * eps (Ne,nPg,ns) = B u_e, u (Ndof) nodal vector, dof of node n comp i = n*dim+i, connect_dof (Ne,nPe) GLOBAL ids. */
int efb_strain(const efb_group* g, const int32_t* connect_dof, const double* u, double* eps, void* stream);

/* Hyperelastic tangent and residual: Operators.NonLinear.SecondPiolaKirchhoffStressTensor and
 * its core __second_piola_block, EasyFEA/FEM/Operators/NonLinear.py:37-201, with HyperElasticState.Compute_De
 * (Models/HyperElastic/_state.py:320-392): F = I + grad u, B = De(u) grad,
 *   Ke (Ne,ndof,ndof) = scale * sum_p wJ (B^T d2Wde B + I (x) dN^T S dN),  Re (Ne,ndof) = scale * sum_p wJ B^T dWde,
 * dWde (Ne,nPg,ns) / d2Wde (Ne,nPg,ns,ns) = the material law's `Compute_dWde/Compute_d2Wde` at the state of u (Kelvin-Mandel),
 * S = the symmetric matrix of dWde; u (Ncoords*dim) nodal, connect_dof GLOBAL node ids; dofs interleaved (x1,y1,z1,x2,...)
 * as the reference returns them.  Ke or Re may be NULL. */
int efb_hyperelastic_Ke_Re(const efb_group* g, const int32_t* connect_dof, const double* u, const double* dWde, const double* d2Wde,
                           double scale, double* Ke, double* Re, void* stream);

/* ---- P2-P7: phase-field law, EasyFEA/Models/_phasefield.py ------------------------------------------- */
enum { EFB_SPLIT_BOURDIN = 0, EFB_SPLIT_AMOR = 1, EFB_SPLIT_MIEHE = 2, EFB_SPLIT_STRESS = 3, EFB_SPLIT_HE = 4 };
enum { EFB_REGU_AT1 = 1, EFB_REGU_AT2 = 2 };

/* host-side parameter block (copied by value at launch) */
typedef struct efb_pf_material {
    int32_t dim;         /* 2 or 3 */
    int32_t split;       /* EFB_SPLIT_* */
    int32_t planeStress; /* 2D only */
    int32_t _pad;
    double E, v, lambda, mu, bulk;    /* Isotropic.get_lambda/get_mu/get_bulk, Models/Elastic/_laws.py:376-405 */
    double C[36];                      /* (ns,ns) row-major, material.C */
    double sqrtC[36], inv_sqrtC[36];   /* Get_sqrt_C_S :223-247 (He split only) */
} efb_pf_material;

/* Calc_C :396-431 and Calc_psi_e_pg :335-358 on a strain field eps (Ne,nPg,ns).  Outputs may be NULL:
 * cP,cM (Ne,nPg,ns,ns); psiP,psiM (Ne,nPg).  If g_e_pg != NULL also writes Cdeg = g*cP + cM (S3,
 * EasyFEA/Simulations/_phasefield.py:462-469).  The 3D repeated-eigenvalue case is chosen per ELEMENT as the
 * reference does (:863,884,904-906); non-finite points are repaired per point (DESIGN.md "degenerate states"). */
int efb_pf_split(const efb_pf_material* m, const double* eps, int64_t Ne, int32_t nPg,
                 int32_t* elem_bits /* (Ne) int32 workspace; may be NULL in 2D or for Bourdin/Amor */, double* cP,
                 double* cM, double* psiP, double* psiM, const double* g_e_pg, double* Cdeg, void* stream);
/* S3 in one pass for one-point simplex elements (TRI3, TETRA4): `PhaseField.__Construct_Elastic_Matrix`,
 * EasyFEA/Simulations/_phasefield.py:444-482 — strain -> split -> C = g(d) cP + cM -> Ke (Ne,ndof,ndof) = scale * wJ B^T C B,
 * u (Nn*dim) / d (Nn) nodal fields, connect_dof GLOBAL node ids.  The strain, g and C(d) arrays of the composition
 * efb_strain -> efb_pf_degradation -> efb_pf_split -> efb_elastic_Ke never exist.  Returns 3 for other element types / quadratures. */
int efb_pf_elastic_Ke(const efb_pf_material* m, const efb_group* g, const int32_t* connect_dof, const double* u, const double* d,
                      double k_res, double scale, double* Ke, void* stream);
/* Get_g_e_pg :295-317: g (Ne,nPg) = (1 - N d_e)^2 + k_res, d (Nn) nodal damage. */
int efb_pf_degradation(const efb_group* g, const int32_t* connect_dof, const double* d, double k_res, double* out,
                       void* stream);
/* history + reaction/source terms: psiP <- max(psiP, psiP_old) (Simulations/_phasefield.py:513-530, in place);
 * r = Get_r_e_pg :253-271, f = Get_f_e_pg :273-293.  psiP_old, r, f may be NULL. */
int efb_pf_history_rf(double* psiP, const double* psiP_old, int64_t n, int regu, double Gc, double l0, double* r,
                      double* f, void* stream);
/* damage sub-problem element system (S4, Simulations/_phasefield.py:540-573), one launch:
 * Ke (Ne,nPe,nPe) = scale * sum_p wJ (r N^T N + k dN^T A dN), Fe (Ne,nPe) = scale * sum_p wJ f N;
 * r,f (Ne,nPg); A (dim,dim) constant. */
int efb_pf_damage_Ke_Fe(const efb_group* g, const double* r, const double* f, const double* A, double k,
                        double scale, double* Ke, double* Fe, void* stream);

/* ---- A1: CSR pattern, EasyFEA/Simulations/_simu.py:1062-1102 ------------------------------------------ */
/* The pattern of a dof_n-block problem is the block expansion of the node adjacency graph, so it is built from
 * node pairs and is bit-exact against scipy's sorted/unique indptr/indices (DESIGN.md "CSR pattern").  The groups
 * that contribute (dict order of Construct_local_matrix_system, _simu.py:1033) are described by HOST arrays of
 * length n_groups (<= 8): device connect pointers (GLOBAL node ids), Ne, nPe.
 * Stages — the caller owns every buffer and reads sizes back between stages:
 *   1. efb_csr_count_node_rows : cnt (Nn) int32 = number of (group, element, local node) touching each node
 *   2. efb_exclusive_scan_i32  : rowptr (Nn+1) int64 from cnt
 *   3. efb_csr_fill_node_rows  : qlist (sum_g Ne*nPe) int64 = ASCENDING element-row ids q = qoff_g + e*nPe_g + a
 *   4. efb_csr_count_adj       : deg (Nn) int32 = number of distinct neighbour nodes (err_flag != 0: > 512)
 *   5. efb_exclusive_scan_i32  : adjptr (Nn+1) int64
 *   6. efb_csr_fill_adj        : adj (nnz_node) int32 ascending neighbour ids per node
 *   7. efb_csr_expand          : indptr (Ndof+1), indices (nnz = dof_n^2 nnz_node) as int32 or int64
 *   8. efb_csr_slot_map        : pos (sum_g Ne*nPe*nPe) int32 = index of node b in adj(node a) for every (e,a,b)
 *   9. efb_csr_inv_map         : the reference's `inv` (int32, k = (g,e,i,j) order, :1098-1101) for API parity
 */
int efb_csr_count_node_rows(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                            const int32_t* nPe_host, int64_t Nn, int32_t* cnt, void* stream);
/* out (n+1) int64 exclusive prefix sums of in (n) int32; workspace >= 8*(n/1024 + 2) bytes */
int efb_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* workspace, void* stream);
int efb_csr_fill_node_rows(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                           const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, int32_t* cursor /* (Nn) scratch */,
                           int64_t* qlist, void* stream);
int efb_csr_count_adj(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                      const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, const int64_t* qlist, int32_t* deg,
                      int32_t* err_flag, void* stream);
int efb_csr_fill_adj(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                     const int32_t* nPe_host, int64_t Nn, const int64_t* rowptr, const int64_t* qlist,
                     const int64_t* adjptr, int32_t* adj, void* stream);
int efb_csr_expand(int64_t Nn, int dof_n, int64_t Ndof, int64_t nnz_node, const int64_t* adjptr, const int32_t* adj,
                   int index_bytes, void* indptr, void* indices, void* stream);
int efb_csr_slot_map(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host,
                     const int32_t* nPe_host, const int64_t* adjptr, const int32_t* adj, int32_t* pos, void* stream);
int efb_csr_inv_map(int n_groups, const int32_t* const* connect_host, const int64_t* Ne_host, const int32_t* nPe_host,
                    int dof_n, const int64_t* adjptr, const int32_t* pos, int32_t* inv, void* stream);
/* vector pattern (isMatrix=False, _simu.py:1075-1078): has (Ndof) int32 = 1 where the row receives an element entry;
 * its exclusive scan is the (Ndof,1) CSR indptr; efb_csr_compact_rows packs a dense vector into that CSR's data */
int efb_csr_row_has_entry(int64_t Nn, int dof_n, int64_t Ndof, const int64_t* rowptr, int32_t* has, void* stream);
int efb_csr_compact_rows(int64_t Ndof, const int64_t* indptr64, const double* dense, double* out, void* stream);

/* ---- A2: deterministic replay, _simu.py:989-1060 ------------------------------------------------------- */
/* data_out (nnz) = np.bincount(inv, weights=concat_g(X_e.ravel())): every slot is summed in ascending entry order k,
 * bit-identical to np.bincount, without atomics.  data_host[g] -> device (Ne_g, ndof_g, ndof_g).  Precondition: the
 * nodes of an element are distinct.  max_deg = max_n deg(n). */
int efb_csr_replay_matrix(int n_groups, const double* const* data_host, const int64_t* Ne_host,
                          const int32_t* nPe_host, int dof_n, int64_t Nn, const int64_t* rowptr, const int64_t* qlist,
                          const int64_t* adjptr, const int32_t* pos, int32_t max_deg, double* data_out, void* stream);
/* out (Nn*dof_n) dense = ordered sum of the element vectors (Ne_g, ndof_g); rows without elements are 0 */
int efb_csr_replay_vector(int n_groups, const double* const* data_host, const int64_t* Ne_host,
                          const int32_t* nPe_host, int dof_n, int64_t Nn, const int64_t* rowptr, const int64_t* qlist,
                          double* out, void* stream);

/* ---- A3 fused: element integration + assembly in one kernel, homogeneous C -------------------------------------
 * `Assembly` (_simu.py:1104-1144) of `LinearizedElasticity(group, C)` (Bilinear.py:62-79) for a homogeneous C, without ever
 * materialising K_e: data_out = CSR data of the rows of the scheduled nodes, summed per slot in ascending element order
 * (deterministic, no atomics; equal to efb_elastic_Ke + efb_csr_replay_matrix up to the association of the products,
 * <= 1e-12 relative).  The schedule is built once per (group, node graph) by the host layer (assembly.FusedSchedule):
 * nodes in clusters of S (spatially compact), cl_nodes (n_clusters*S, 4) int64 = {node id or -1, dim*dim*adjptr[n],
 * deg | cnt << 32, first task}, cl_ne (n_clusters) / cl_conn (n_clusters, cap_e, nPe) = coordinate rows of every element
 * touching a cluster (-1 = empty), desc (n_tasks) = element index inside its cluster | local node << 16, tpos (n_tasks, nPe) = slot of every
 * column node, cap_e = max elements of a cluster, max_deg = max neighbours of a node.  C_host (ns,ns) and w_pg_host (nPg)
 * are HOST arrays.  Returns 3 (and sets the error string) when the configuration is outside this kernel (element type,
 * shared-memory budget, non-positive quadrature weights): the caller then uses the two-kernel path. */
int efb_assemble_elastic(const efb_group* g, const double* C_host, const double* w_pg_host, double scale, int n_clusters, int S,
                         int cap_e, int max_deg, const int64_t* cl_nodes, const int32_t* cl_ne, const int32_t* cl_conn,
                         const int32_t* desc, const int32_t* tpos, double* out, void* stream);
/* dynamic shared memory (bytes) the fused kernel needs for this configuration, or -1 when there is no instantiation */
int efb_assemble_elastic_smem(int dim, int nPe, int nPg, int S, int cap_e, int max_deg);
/* nodes per warp of the instantiation that serves (dim, nPe, nPg): S must be a multiple of it and at most 16 times it; -1
 * when there is no instantiation */
int efb_assemble_elastic_group(int dim, int nPe, int nPg);

/* ---- A3 fused, HEXA8 with 8 Gauss points on the FP64 tensor-core instruction (DMMA.8x8x4) ------------------------
 * Same contract as efb_assemble_elastic (`Assembly` of `LinearizedElasticity(group, C)`, _simu.py:1104-1144 with
 * Bilinear.py:62-79, homogeneous C, K_e never materialised, every CSR slot summed in ascending element order by one lane, no
 * atomics) for clusters of 16 nodes: the element contraction T_e = G^T G runs as m8n8k4 FP64 MMAs, the T rows of the owned
 * nodes are staged in shared memory and a per-cluster gather program (built once by assembly.MmaSchedule) sums them per CSR
 * block.  recs (n_clusters, rec_words) int32 = one record per cluster, streamed into shared memory by bulk copies:
 *   conn (cap4, 8) coordinate rows of the cluster's elements, -1 = empty, cap4 a multiple of 4 |
 *   rowslot (cap4, 8) int16 = staging slot | (node-in-cluster & 7) << 8 of (element, local node), -1 = node not owned |
 *   nodes (16, 2) int64 = {dim*dim*adjptr[n] (-1: padding), deg} |
 *   hdr (1 + rmax) = rounds R of the cluster, then per round (word offset inside the cluster program << 8) | trips;
 * rec_words a multiple of 4; t_cap = max tasks of a cluster (<= 256); prog_off (n_clusters + 1) int64 = offsets in int32 words
 * of the clusters' programs inside `prog`, pw_max = the longest program; a round of c trips is [dest of the 32 lanes =
 * node-in-cluster << 16 | slot, -1 = none][32 x cpad uint16 sources = staging offset in doubles of the contribution,
 * t_cap*72 (a block of zeros) for lanes without one], cpad = c rounded up to 1, 2, 4 or a multiple of 8.  Returns 3 when the configuration is outside the kernel. */
int efb_assemble_elastic_mma(const efb_group* g, const double* C_host, const double* w_pg_host, double scale, int n_clusters,
                             int cap4, int t_cap, int rmax, int rec_words, int pw_max, const int32_t* recs,
                             const int64_t* prog_off, const int32_t* prog, double* out, void* stream);
/* dynamic shared memory (bytes) of the CTA of efb_assemble_elastic_mma for this configuration */
int efb_assemble_elastic_mma_smem(int t_cap, int rec_words, int pw_max);

/* ---- consumer: Jacobi-PCG building blocks (north star; the reference dispatches in Solvers.py:225-394) ---- */
/* All scalars stay on the device; dot products are fixed-order two-stage reductions: producers write
 * efb_pcg_partials_size() doubles of partials, efb_pcg_reduce folds them (deterministic). */
int efb_pcg_partials_size(void);
/* y[r] = sum_k data[k] x[indices[k]] for local rows r < nrows (masked rows give 0); x is indexed by GLOBAL column;
 * if dot_partials != NULL also the partials of sum_r x[x_row_offset + r] * y[r].  lanes_per_row in {4,8,16,32} (node-block form: {2,4,8,16,32}). */
int efb_spmv_csr(int64_t nrows, int index_bytes, const void* indptr, const void* indices, const double* data,
                 const double* x, int64_t x_row_offset, const uint8_t* row_mask, double* y, double* dot_partials,
                 int lanes_per_row, void* stream);
/* same product for a matrix assembled by efb_csr_replay_matrix, read through the node adjacency it was built from
 * (adjptr/adj of the pattern): y = A x over the dof rows of nodes [0, n_nodes); x, row_mask, dot_partials as above */
int efb_spmv_nodeblock(int64_t n_nodes, int dof_n, const int64_t* adjptr, const int32_t* adj, const double* data,
                       const double* x, int64_t x_row_offset, const uint8_t* row_mask, double* y, double* dot_partials,
                       int lanes_per_node, void* stream);
int efb_csr_diagonal(int64_t nrows, int64_t row_offset, int index_bytes, const void* indptr, const void* indices,
                     const double* data, double* diag, void* stream);
int efb_pcg_inv_diag(int64_t n, const double* diag, const uint8_t* free_mask, double* out, void* stream);
int efb_pcg_dot(int64_t n, const double* a, const double* b, double* partials, void* stream);
int efb_pcg_reduce(const double* partials, int m /* 1 or 2 */, double* out, void* stream);
/* r = mask(b - Ax); z = r/diag; p = z; partials of (r.z, r.r) */
int efb_pcg_init(int64_t n, const double* b, const double* Ax, const double* inv_diag, const uint8_t* free_mask, double* r,
                 double* z, double* p, double* partials, void* stream);
/* alpha = rz/pAp; x += alpha p; r -= alpha Ap; z = r/diag; partials of (r.z, r.r) */
int efb_pcg_update_xr(int64_t n, const double* rz, const double* pAp, const double* p, const double* Ap, double* x,
                      double* r, const double* inv_diag, const uint8_t* free_mask, double* z, double* partials,
                      void* stream);
/* p = z + (rz_new/rz_old) p */
int efb_pcg_update_p(int64_t n, const double* rz_new, const double* rz_old, const double* z, const uint8_t* free_mask,
                     double* p, void* stream);

/* ---- fused Jacobi-PCG iterations with peer-memory communication ----------------------
 * Three kernels per iteration and no host round trip:  (1) Ap = A p with the p.Ap partials, (2) x, r, z update with the
 * (r.z, r.r) partials, (3) p' = z + beta p written to the OTHER p buffer.  The last CTA of a kernel folds the per-CTA
 * partials in fixed order and STORES the result into the control block of every rank (peer memory over NVLink);
 * the next kernel's CTAs wait for the `world` contributions and add them in rank order, so every rank obtains the same
 * bits without a collective call.  Kernel (3) also stores the interface entries of p' straight into the halo segments
 * of the neighbours' p buffers and raises their halo flags; kernel (1) waits for its neighbours' flags.
 * Replaces, for row-sharded runs, the per-iteration NCCL send/recv + 2 all-reduces (the reference: PETSc KSP over
 * mpi4py, Simulations/Solvers.py:732-847).  world == 1 runs the same kernels on local memory. */
#define EFB_MAX_RANKS 8
typedef struct efb_pcg_system {
    int64_t nrows;            /* owned dof rows */
    int32_t kind;             /* 0: CSR (indptr/indices, index_bytes), 1: node blocks (adjptr/adj, dof_n) */
    int32_t index_bytes;      /* 4 or 8 (kind 0) */
    int32_t dof_n;            /* 1..3 (kind 1) */
    int32_t lanes;            /* lanes per row / node: 4, 8, 16 or 32 */
    const void* indptr;       /* kind 0: (nrows+1); kind 1: adjptr (n_nodes+1) int64 */
    const void* indices;      /* kind 0: (nnz);     kind 1: adj int32 */
    const double* data;
    const uint8_t* free_mask; /* (nrows) or NULL */
    const double* inv_diag;   /* (nrows), 0 on constrained rows */
    double* x;                /* (nrows) */
    double* r;
    double* z;
    double* Ap;
    double* partials;         /* efb_pcg_partials_size() doubles */
    double* s;                /* (nrows) extra vector of the single-reduction form (efb_pcg_iterate_cg2), else NULL */
} efb_pcg_system;

typedef struct efb_pcg_peer {
    int32_t world, rank;
    int32_t n_send, n_recv;                   /* neighbour counts of the halo exchange */
    int32_t send_rank[EFB_MAX_RANKS];
    int32_t recv_rank[EFB_MAX_RANKS];
    int64_t send_ptr[EFB_MAX_RANKS + 1];      /* send_idx[send_ptr[i] : send_ptr[i+1]] goes to send_rank[i] */
    int64_t send_dst[EFB_MAX_RANKS];          /* first entry of this rank's segment inside that neighbour's p buffers */
    void* base[EFB_MAX_RANKS];                /* region of every rank as mapped HERE (own region at [rank]) */
    int64_t pbuf_off[EFB_MAX_RANKS][4];       /* byte offsets of the two p buffers and the two z buffers (Chebyshev form) inside each region */
    const int32_t* send_idx;                  /* owned local dof ids to push, concatenated per neighbour */
    uint64_t ar_seq;                          /* reductions already published on this communicator */
    uint64_t halo_seq;                        /* halo pushes already published */
    /* the same push plan seen from the rows (single-reduction form: the CTA that updates a row also stores it into the
     * neighbours, no separate push kernel): push_id (nrows) = -1 or c; entries push_ptr[c] .. push_ptr[c+1] of
     * push_nbr (index into send_rank) / push_pos (entry inside that neighbour's vector).  NULL when n_send == 0. */
    const int32_t* push_id;
    const int64_t* push_ptr;
    const int32_t* push_nbr;
    const int64_t* push_pos;
} efb_pcg_peer;

/* a region starts with a control block of efb_pcg_ctrl_bytes() bytes (zero it once); offsets of the device scalars
 * inside it: out[0] = rz[2] ping-pong and out[1] = rr (in doubles), out[2] = error flag (uint32 index),
 * out[3] = iterations run by the last persistent launch (uint64 index), out[4] = (r.z, z.Az, r.r) of the start vector for
 * the single-reduction form (in doubles) */
int efb_pcg_ctrl_bytes(void);
int efb_pcg_ctrl_layout(int32_t* out5);
/* enqueue n_iters iterations starting at iteration `it0` (p_it lives in p buffer it & 1, r.z of it in rz[it & 1]).
 * The caller advances peer->ar_seq by 2*n_iters and halo_seq by n_iters afterwards. */
int efb_pcg_iterate(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, void* stream);
/* The same iterations with a Chebyshev-Jacobi polynomial preconditioner: z = q(D^-1 A) D^-1 r, q the degree `degree`-1
 * Chebyshev polynomial of [lmin, lmax] (bounds of the spectrum of D^-1 A the polynomial damps; lmax must not be below the
 * largest eigenvalue).  `degree` - 1 more products per iteration (each with one neighbour halo flag, no all-reduce), about
 * `degree` times fewer iterations: what strong-scaled shards need, since an iteration costs two all-reduces whatever the
 * shard size.  d_vec (nrows) is scratch; z lives in peer buffers 2 and 3 of efb_pcg_peer (with halos), p as in
 * efb_pcg_iterate.  Before iteration 0 the caller puts p_0 = z_0 = q(D^-1 A) D^-1 r_0 (owned + halo) in p buffer 0 and r.z in
 * the control block.  The caller advances ar_seq by 2*n_iters and halo_seq by degree*n_iters afterwards. */
int efb_pcg_iterate_cheb(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, int degree, double lmin,
                         double lmax, double* d_vec, const float* data32, int cheb_lanes, void* stream);
/* data32 (same layout as sys->data, or NULL): the matrix values in single precision for the products INSIDE the polynomial — half
 * the matrix traffic of degree - 1 of the degree products of an iteration.  A polynomial in fl32(A) is still a fixed symmetric
 * operator, hence a valid preconditioner; the outer product, residuals, vectors and dot products stay FP64.  Those products use
 * the block form of the node product (a lane takes whole neighbour blocks) with cheb_lanes lanes per node (0 = 4). */
int efb_cast_f32(int64_t n, const double* src, float* dst, void* stream);
/* unfused building block of the same recurrence: d = c1 d + c2 D^-1 (r - t), z += d; t = A z of the caller (NULL: 0) */
int efb_pcg_cheb_update(int64_t n, const double* r, const double* t, const double* inv_diag, double c1, double c2, double* d, double* z,
                        void* stream);
/* Single-reduction form (Chronopoulos-Gear): the same iterates with ONE all-reduce per iteration — two kernels (vector
 * update, SpMV) and two cross-GPU sync points per iteration instead of three and three (the interface entries of z are stored into the neighbours by
 * the update kernel itself, through the row-wise push plan of efb_pcg_peer).  z lives in the two peer buffers
 * (its halo is what travels), p in sys->z, w = A z in sys->Ap, s = A p in sys->s.  Before iteration 0 the caller puts z_0
 * (owned + halo) in buffer 0, w_0 = A z_0 in sys->Ap, zeros in p and s, and (r.z, z.Az, r.r) in the control block
 * (layout out[4]).  The caller advances ar_seq and halo_seq by n_iters afterwards; r.r of the last iterate is in the
 * control block when the call's work is done. */
int efb_pcg_iterate_cg2(const efb_pcg_system* sys, const efb_pcg_peer* peer, int n_iters, int64_t it0, void* stream);
/* Persistent form for node-block systems (kind 1): ONE cooperative kernel runs iterations it0, it0+1, ... until
 * r.r <= target_rr or max_iters are done; the three steps are separated by grid barriers, every CTA folds the partials
 * itself, and every CTA of every rank leaves in the same iteration (same bits everywhere) — no host round trip and no
 * launch per iteration.  sys->partials holds efb_pcg_partials_size() doubles.  Afterwards the control block holds the
 * iteration count (layout out[3]) and r.r; the caller advances ar_seq by 2*iterations and halo_seq by iterations. */
int efb_pcg_solve_persistent(const efb_pcg_system* sys, const efb_pcg_peer* peer, int64_t it0, int64_t max_iters,
                             double target_rr, void* stream);

/* peer-mappable device memory (plain cudaMalloc so that cudaIpc can export it) */
int efb_peer_alloc(int64_t bytes, void** ptr);                 /* zero-filled */
int efb_peer_free(void* ptr);
int efb_peer_export(void* ptr, void* handle64 /* 64 bytes out */);
int efb_peer_open(const void* handle64, void** ptr);           /* maps a region exported by another process */
int efb_peer_close(void* ptr);

/* ---- time-scheme system build -------------------------------------------------
 * out[i] = sum_{k<m} coefs[k] * vecs[k][i], m <= 4; coefs and the pointer table are HOST arrays, vecs[k] and out device
 * arrays of n doubles (out may be one of the inputs).  Serves A = coefK K + coefC C + coefM M on the shared CSR pattern
 * (_simu.py:1890-1894), the history terms of the right-hand side (:1777-1853) and the correctors (:1552-1657). */
int efb_lincomb(int64_t n, int m, const double* coefs, const double* const* vecs_host, double* out, void* stream);

/* ---- post-processing fields ----------------------------------------------------
 * sigma (Ne,nPg,ns) = C eps, `Calc_Sigma_e_pg` (Models/Elastic/_laws.py:159-185); C_mode 0: (ns,ns), 1: (Ne,ns,ns),
 * 2: (Ne,nPg,ns,ns) */
int efb_hooke(int64_t Ne, int nPg, int ns, const double* eps, const double* C, int C_mode, double* sigma, void* stream);
/* per-element result of a Kelvin-Mandel strain/stress field, `Result_strain_or_stress_field_e` (Models/_utils.py:302-430):
 * shear components / coef, then the mean over the Gauss points of component `what` (>= 0), of the von Mises value (-1),
 * or of every component (-2: out is (Ne, ns)) */
int efb_field_result(int64_t Ne, int nPg, int dim, const double* field, int what, double coef, double* out, void* stream);
/* Wdef_e (Ne) = scale * sum_p wJ * 1/2 sigma.eps, `_Calc_Psi_Elas` (Simulations/_elastic.py:323-396) */
int efb_energy_e(int64_t Ne, int nPg, int ns, const double* eps, const double* sigma, const double* wJ, double scale,
                 double* out, void* stream);
/* out (Nn, ncols) = element values averaged over the elements around each node, `Mesh.Get_Node_Values`
 * (FEM/_mesh.py:822-873); rowptr/qlist of the assembly pattern (efb_csr_fill_node_rows), one group of nPe-node elements */
int efb_node_values(int64_t Nn, const int64_t* rowptr, const int64_t* qlist, int nPe, const double* values_e, int ncols,
                    double* out, void* stream);

/* send-buffer packing of the PCG halo exchange (row-sharded runs): dst[i] = src[idx[i]] */
int efb_pack_f64(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EASYFEA_B200_H */
