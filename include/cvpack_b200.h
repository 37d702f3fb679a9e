/*
 * cvpack_b200 -- C ABI of the B200-native batched collective-variable evaluator.
 *
 * The reference (RedesignScience/cvpack) has no FFI of its own: each CV is an OpenMM Force and the
 * boundary under its Python API is `openmm.Context.getState(getEnergy|getForces, groups=1<<g)`
 * (cvpack/utils.py:140) -- "energy" is the CV value, "forces" are minus its gradient, one frame per
 * call.  This header is what a binding for that path would bind instead: a CV is compiled once
 * into an opaque kernel spec (`cvp_cv`, the analogue of the OpenMM Force object the cvpack
 * constructor configures) and then evaluated for a whole batch of frames per call.
 *
 * Conventions
 *   - coordinates: nm, row-major [B][N][3], float (CVP_F32) or double (CVP_F64)
 *   - gradient: +dCV/dr, same layout and type as the coordinates (OpenMM forces = -gradient)
 *   - box: three row vectors a,b,c as 9 numbers (same type as coordinates); CVP_BOX_SHARED = one
 *     box [9] for every frame, CVP_BOX_PER_FRAME = [B][9].  Must be in OpenMM's reduced form
 *     (a=(ax,0,0), b=(bx,by,0), c=(cx,cy,cz)).  Set CVP_FLAG_TRICLINIC if any off-diagonal != 0.
 *   - atom indices: int32, 0-based, < num_atoms; index handling is exact (no float round trips)
 *   - every function returns CVP_OK (0) or a negative cvp_status; nothing throws across the ABI;
 *     cvp_last_error() gives the message of the last failure on the calling thread
 *   - the caller owns every buffer; the library owns only spec-internal device copies of index /
 *     parameter arrays and a per-spec workspace, both released by cvp_destroy
 *   - cvp_evaluate is asynchronous with respect to the host and ordered on `stream`; a spec must be
 *     evaluated on ONE stream / thread at a time: its workspace (and the staging buffers of the host
 *     path) are shared by all calls on that spec.  Build one spec per stream for concurrency.
 *   - alignment: any element-aligned pointer is accepted.  When `positions` and `grad` are both
 *     16-byte aligned (every cudaMalloc / torch allocation is) and the group is a contiguous run
 *     of atoms starting at a multiple of 4 in frames of a multiple of 4 atoms, RMSD and radius of
 *     gyration take the vectorised / TMA streaming kernels; otherwise the generic kernels run --
 *     same results to rounding, never a misaligned access.
 *   - entry points taking a device ordinal (`*_host`, cvp_measure_fp32_peak) switch to that device
 *     for the duration of the call and restore the caller's current device before returning
 */
#ifndef CVPACK_B200_H
#define CVPACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cvp_cv cvp_cv;

typedef enum {
  CVP_OK = 0,
  CVP_ERR_INVALID = -1,     /* bad argument (null pointer, index out of range, size mismatch)   */
  CVP_ERR_CUDA = -2,        /* a CUDA runtime call failed; message has the CUDA error string      */
  CVP_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve                              */
  CVP_ERR_NO_DEVICE = -4    /* no usable sm_100 device                                            */
} cvp_status;

enum { CVP_F32 = 0, CVP_F64 = 1 };
enum { CVP_BOX_NONE = 0, CVP_BOX_SHARED = 1, CVP_BOX_PER_FRAME = 2 };
enum {
  CVP_FLAG_TRICLINIC = 1,   /* box has non-zero off-diagonal elements                             */
  CVP_FLAG_ACCUMULATE = 2   /* add into value/grad instead of overwriting them                    */
};

/* Step functions S(x) used by contact-type sums (the whitelisted forms of cvpack's `stepFunction`
 * strings; number_of_contacts.py:132, residue_coordination.py:113, helix_rmsd_content.py:135). */
enum {
  CVP_STEP_RATIONAL = 0,    /* 1/(1+x^n), n = p0 (any positive integer)                           */
  CVP_STEP_HEAVISIDE = 1,   /* step(1-x): 1 if x <= 1 else 0; zero gradient                       */
  CVP_STEP_RATIONAL_PAIR = 2, /* (1+x^n)/(1+x^n+x^(2n)), n = p0                                   */
  CVP_STEP_PLUMED = 3       /* (1-x^n)/(1-x^m), n = p0, m = p1; n/m at x == 1                      */
};

/* ---- pair sums: NumberOfContacts, AttractionStrength, ShortestDistance ---------------------- */
/* Replaces an OpenMM CustomNonbondedForce with one interaction group (group1, group2):
 * number_of_contacts.py:143-161, attraction_strength.py:189-231, shortest_distance.py:122-138.
 * The sum runs over unordered pairs {i,j}, i in group1, j in group2, i != j, each pair once even
 * if both atoms are in both groups, excluding `exclusions`, and only r < cutoff.                 */
enum {
  CVP_PAIR_CONTACTS = 0,    /* scale * sum S(r/r0) * sw(r)                                        */
  CVP_PAIR_ATTRACTION = 1,  /* scale * sum -sgn_i sgn_j [4 eps (1/y^2-1/y) + ke min(0,qiqj)(1/x+(x^2-3)/2)] sw(r) */
  CVP_PAIR_SOFTMIN = 2      /* (rc e^-beta + sum r w)/(e^-beta + sum w), w = exp(-beta r/rc)       */
};

typedef struct {
  int32_t kind;
  int32_t num_atoms;
  int32_t n1;
  const int32_t* group1;
  int32_t n2;
  const int32_t* group2;
  int32_t num_exclusions;
  const int32_t* exclusions;   /* [num_exclusions][2] atom pairs never counted                    */
  int32_t periodic;            /* minimum-image convention on/off                                 */
  double cutoff;               /* r >= cutoff skipped; <= 0 means no cutoff                       */
  double switch_distance;      /* sw(r) = 1 - 10t^3 + 15t^4 - 6t^5 for r > it; <= 0 means none    */
  double scale;                /* 1/reference                                                     */
  int32_t step_kind;           /* CVP_PAIR_CONTACTS                                               */
  double step_p0, step_p1;
  double r0;                   /* threshold distance of the step function                         */
  const double* charge;        /* CVP_PAIR_ATTRACTION: per atom [num_atoms]                       */
  const double* sigma;
  const double* epsilon;
  const double* sign;          /* may be NULL (all +1)                                            */
  double coulomb_constant;     /* 138.93545764438198 kJ nm /(mol e^2)                             */
  double beta;                 /* CVP_PAIR_SOFTMIN                                                */
} cvp_pair_sum_desc;

int cvp_pair_sum_create(const cvp_pair_sum_desc* desc, cvp_cv** out);

/* ---- centroid pair sums: ResidueCoordination -------------------------------------------------- */
/* Replaces CustomCentroidBondForce(2, ...) with one bond per (residue of set 1, residue of set 2):
 * residue_coordination.py:129-144.  value = scale * sum_{I<n1} sum_{J<n2} S(|R_J - R_I|/r0),
 * R = sum_a w_a r_a over raw coordinates, minimum image on R_J - R_I only.                        */
typedef struct {
  int32_t num_atoms;
  int32_t n1, n2;                 /* number of centroid groups in each set                        */
  const int32_t* group_offsets;   /* [n1+n2+1] CSR offsets into group_atoms / group_weights        */
  const int32_t* group_atoms;
  const double* group_weights;    /* normalised within each group                                 */
  int32_t periodic;
  int32_t step_kind;
  double step_p0, step_p1;
  double r0;
  double scale;
} cvp_centroid_pair_sum_desc;

int cvp_centroid_pair_sum_create(const cvp_centroid_pair_sum_desc* desc, cvp_cv** out);

/* ---- RMSD family: RMSD, CompositeRMSD, Helix/SheetRMSDContent, PathInRMSDSpace ---------------- */
/* Replaces openmm.RMSDForce (rmsd.py:96), openmmcppforces.CompositeRMSDForce (composite_rmsd.py:
 * 135-137) and the CustomCVForce compositions over RMSD CVs (base_rmsd_content.py:44-79,
 * base_path_cv.py:43-62).  Each *group* is one optimal-superposition problem; each group may be
 * split in *bodies* centred separately but rotated together (composite RMSD).                     */
enum {
  CVP_COMBINE_IDENTITY = 0,       /* value = rmsd of the single group                             */
  CVP_COMBINE_STEP_SUM = 1,       /* value = scale * sum_g S(rmsd_g/r0)                           */
  CVP_COMBINE_PATH_PROGRESS = 2,  /* s = sum_g g w_g /((G-1) sum w),  w_g = exp(-rmsd_g^2/(2 sigma^2)) */
  CVP_COMBINE_PATH_DEVIATION = 3  /* z = -2 sigma^2 ln sum_g w_g                                  */
};

typedef struct {
  int32_t num_atoms;
  int32_t num_groups;
  const int32_t* group_offsets;   /* [num_groups+1] CSR offsets into atoms / body_ids / reference  */
  const int32_t* atoms;           /* atom index of each entry; order pairs with reference rows     */
  const int32_t* body_ids;        /* body of each entry within its group, 0-based; NULL = all 0    */
  const double* reference;        /* [entries][3] reference coordinates (any origin)               */
  int32_t combine;
  int32_t step_kind;
  double step_p0, step_p1;
  double r0;
  double scale;
  double sigma;
} cvp_rmsd_desc;

int cvp_rmsd_create(const cvp_rmsd_desc* desc, cvp_cv** out);

/* ---- radius of gyration: RadiusOfGyration, RadiusOfGyrationSq --------------------------------- */
/* Replaces the CustomCentroidBondForce constructions of radius_of_gyration.py:85-92 and
 * radius_of_gyration_sq.py:87-89: rg^2 = (1/n) sum_i |r_i - c|^2, c = sum_i w_i r_i.            */
typedef struct {
  int32_t num_atoms;
  int32_t n;
  const int32_t* group;
  const double* weights;          /* [n] weights of the centre, normalised (1/n or m_i/M)          */
  int32_t periodic;               /* minimum image on each r_i - c                                 */
  int32_t squared;                /* 1: value = rg^2 (nm^2); 0: value = rg (nm)                     */
} cvp_gyration_desc;

int cvp_gyration_create(const cvp_gyration_desc* desc, cvp_cv** out);

/* ---- bonded terms: Distance, Angle, Torsion, Helix{Torsion,Angle,HBond}Content,
 *      TorsionSimilarity --------------------------------------------------------------------------- */
/* Replaces CustomBondForce / CustomAngleForce / CustomTorsionForce with one expression for all
 * terms (distance.py:57-59, angle.py:73-75, torsion.py:80-82, helix_torsion_content.py:126-145,
 * helix_angle_content.py:118-128, helix_hbond_content.py:104-114).
 * value = sum_t f(q_t): q = |r2-r1| (2 atoms), acos angle at atom 2 (3 atoms), the dihedral
 * (4 atoms, sign convention of torsion.py:22-33), or the difference of two dihedrals (8 atoms:
 * the CustomCompoundBondForce of torsion_similarity.py:89-93).                                     */
enum {
  CVP_TERM_IDENTITY = 0,          /* f(q) = coefficient * q                                        */
  CVP_TERM_BOXCAR = 1,            /* f(q) = coefficient / (1 + x^(2m)), x = (q - ref_t)/tolerance   */
  CVP_TERM_HALF_COSINE = 2        /* f(q) = coefficient * (1 + cos q)/2; 8-atom terms only          */
};

typedef struct {
  int32_t num_atoms;
  int32_t num_terms;
  int32_t atoms_per_term;         /* 2, 3, 4 or 8                                                  */
  const int32_t* atoms;           /* [num_terms][atoms_per_term]                                   */
  int32_t function;
  const double* reference;        /* [num_terms] (CVP_TERM_BOXCAR), may be NULL otherwise          */
  double tolerance;
  int32_t half_exponent;          /* m                                                             */
  double coefficient;
  int32_t periodic;
} cvp_bonded_desc;

int cvp_bonded_create(const cvp_bonded_desc* desc, cvp_cv** out);

/* ---- lifetime, evaluation ---------------------------------------------------------------------- */
int cvp_destroy(cvp_cv* cv);

/* Number of atoms that can carry a non-zero gradient, and their sorted indices (for compact
 * gathers across GPUs).  `atoms` may be NULL to query the count only.                             */
int cvp_active_atoms(const cvp_cv* cv, int32_t* count, int32_t* atoms);

/* Cutoff (nm) of a spec whose sum uses the minimum-image convention, 0 if it has none.  Like OpenMM
 * ("The cutoff distance cannot be greater than half the periodic box size"), such a spec must only
 * be evaluated in boxes whose smallest edge is at least twice this number: cvp_evaluate_host
 * checks its (host) boxes and returns CVP_ERR_INVALID; for cvp_evaluate the boxes are device data
 * and the check is the binding's (the Python host does it from its host copy of the box).        */
int cvp_periodic_cutoff(const cvp_cv* cv, double* cutoff);

/* Evaluate on DEVICE buffers, asynchronously on `stream` (a cudaStream_t; NULL = default stream).
 * `value` [B] and `grad` [B][N][3] have the coordinate type; `grad` may be NULL (value only).
 * Replaces Context.getState(getEnergy=True, getForces=True, groups=1<<g) per frame
 * (cvpack/utils.py:140).                                                                          */
int cvp_evaluate(cvp_cv* cv, int dtype, const void* positions, const void* box, int box_mode,
                 int64_t num_frames, int64_t num_atoms, void* value, void* grad, int flags,
                 void* stream);

/* Same through HOST buffers (pinned for full speed): frames are cut in chunks (a short first and
 * last one) that flow through host->device copy, kernels and device->host copy on the streams of
 * two pipeline slots -- four slots and two workspaces for pair-sum specs, whose consecutive chunks
 * may run at once.  Staging buffers and workspaces belong to the spec and live until cvp_destroy.
 * Synchronous.                                                                                   */
int cvp_evaluate_host(cvp_cv* cv, int dtype, const void* positions, const void* box, int box_mode,
                      int64_t num_frames, int64_t num_atoms, void* value, void* grad, int flags,
                      int device);

/* Effective mass 1/sum_i |g_i|^2/m_i over atoms with non-zero gradient, +inf if none
 * (cvpack/utils.py:172-184).  `grad` [B][N][3] device, `masses` [N] device double, `emass` [B]
 * device, coordinate type.                                                                        */
int cvp_effective_mass(int dtype, const void* grad, const double* masses, int64_t num_frames,
                       int64_t num_atoms, void* emass, void* stream);

/* Value and effective mass in one call, without handing [B][N][3] to the caller: the gradient of a
 * slab of frames stays in a spec-owned scratch buffer between the CV kernels and the reduction.
 * Replaces CollectiveVariable.getEffectiveMass (collective_variable.py:237-309 ->
 * utils.compute_effective_mass, utils.py:151-184).  Device buffers; `masses` [N] device double.   */
int cvp_evaluate_with_mass(cvp_cv* cv, int dtype, const void* positions, const void* box,
                           int box_mode, int64_t num_frames, int64_t num_atoms,
                           const double* masses, void* value, void* emass, int flags,
                           void* stream);

/* Same through HOST buffers (`masses` [N] host double): only [B] values and [B] masses travel back;
 * `grad` (host, [B][N][3]) may be NULL -- pass a buffer to get the gradient as well.  Synchronous. */
int cvp_evaluate_with_mass_host(cvp_cv* cv, int dtype, const void* positions, const void* box,
                                int box_mode, int64_t num_frames, int64_t num_atoms,
                                const double* masses, void* value, void* emass, void* grad,
                                int flags, int device);

/* ---- PathInCVSpace: a path of milestones in the space of other collective variables ----------- */
/* Replaces the CustomCVForce expression assembled by path_in_cv_space.py:141-158 and
 * base_path_cv.py:43-62.  The inner CVs are evaluated with cvp_evaluate; this entry combines their
 * values [num_variables][B] (device, coordinate type) into the path value [B] and the partial
 * derivatives d value / d cv_j [num_variables][B] (`out_coef`, may be NULL); cvp_axpy_frames then
 * accumulates coef_j[b] * gradient_j[b][N][3] into the gradient of the path variable.             */
typedef struct {
  int32_t num_variables;
  int32_t num_milestones;         /* >= 2                                                          */
  const double* milestones;       /* [num_milestones][num_variables], MD units                     */
  const double* periods;          /* [num_variables]; 0 = not periodic, else upper - lower bound   */
  const double* scales;           /* [num_variables] characteristic scales (> 0)                   */
  double sigma;                   /* width of the Gaussian kernels                                 */
  int32_t metric;                 /* CVP_COMBINE_PATH_PROGRESS or CVP_COMBINE_PATH_DEVIATION       */
} cvp_path_cv_desc;

int cvp_path_cv_combine(const cvp_path_cv_desc* desc, int dtype, const void* values,
                        int64_t num_frames, void* out_value, void* out_coef, void* stream);

/* ---- MetaCollectiveVariable: a function of other collective variables and named parameters ---- */
/* Replaces the CustomCVForce of meta_collective_variable.py:114-141, whose energy function is a
 * user-supplied Lepton expression of the inner variables' names and of named global parameters.
 * The binding compiles that string into a postfix program (cvpack_b200/expression.py does, with
 * Lepton's grammar); cvp_expr_eval interprets it for every frame: `values` [num_variables][B]
 * (device, coordinate type) -> `out_value` [B] and, if not NULL, `out_partial`
 * [num_variables + num_parameters][B] = d value / d input (inner variables first, then the
 * parameters: meta_collective_variable.py:236-261 getParameterDerivatives).  cvp_axpy_frames then
 * assembles the gradient.  Derivative rules at kinks are Lepton's (see csrc/expr.cu).           */
enum { CVP_EXPR_MAX_OPS = 192, CVP_EXPR_MAX_INPUTS = 16, CVP_EXPR_MAX_TEMPS = 16 };
enum {
  CVP_EXPR_CONST = 0,   /* push `value`                                                            */
  CVP_EXPR_INPUT = 1,   /* push input `arg` (inner variable, or parameter arg - num_variables)      */
  CVP_EXPR_LOAD = 2,    /* push temporary slot `arg` (a `name = expression` definition)             */
  CVP_EXPR_STORE = 3,   /* pop into temporary slot `arg`                                            */
  CVP_EXPR_ADD = 10, CVP_EXPR_SUB, CVP_EXPR_MUL, CVP_EXPR_DIV, CVP_EXPR_POW, CVP_EXPR_MIN,
  CVP_EXPR_MAX, CVP_EXPR_ATAN2,                                  /* binary: pop b, pop a, push a.b */
  CVP_EXPR_SELECT = 20,                                          /* pop no, yes, cond              */
  CVP_EXPR_NEG = 30, CVP_EXPR_SQRT, CVP_EXPR_EXP, CVP_EXPR_LOG, CVP_EXPR_SIN, CVP_EXPR_COS,
  CVP_EXPR_TAN, CVP_EXPR_SEC, CVP_EXPR_CSC, CVP_EXPR_COT, CVP_EXPR_ASIN, CVP_EXPR_ACOS,
  CVP_EXPR_ATAN, CVP_EXPR_SINH, CVP_EXPR_COSH, CVP_EXPR_TANH, CVP_EXPR_ERF, CVP_EXPR_ERFC,
  CVP_EXPR_STEP, CVP_EXPR_DELTA, CVP_EXPR_SQUARE, CVP_EXPR_CUBE, CVP_EXPR_RECIP, CVP_EXPR_ABS,
  CVP_EXPR_FLOOR, CVP_EXPR_CEIL                                  /* unary                          */
};
typedef struct {
  int32_t code;
  int32_t arg;
  double value;
} cvp_expr_op;
typedef struct {
  int32_t num_ops;
  int32_t num_variables;
  int32_t num_parameters;
  int32_t reserved;
  double parameters[CVP_EXPR_MAX_INPUTS];
  cvp_expr_op ops[CVP_EXPR_MAX_OPS];
} cvp_expr_program;

int cvp_expr_eval(const cvp_expr_program* program, int dtype, const void* values,
                  int64_t num_frames, void* out_value, void* out_partial, void* stream);

/* y[b][N][3] = (accumulate ? y : 0) + coef[b] * x[b][N][3]; device buffers of the given dtype.   */
int cvp_axpy_frames(int dtype, const void* coef, const void* x, void* y, int64_t num_frames,
                    int64_t num_atoms, int accumulate, void* stream);

/* ---- gathering gradients across the GPUs of one box (NVLink / NVSwitch peer memory) ------------ */
/* north_star: "NCCL over NVLink used only to gather CV values and gradients"; SURVEY section 8(b) calls
 * the entry cvp_allgather.  Values ([B] numbers) are gathered by the host language over NCCL; for
 * gradients the library has its own collective: every rank owns a gather buffer [B_total][N][3]
 * exported as a 64-byte CUDA IPC handle, ranks exchange the handles out of band and map each
 * other's buffers, and cvp_peer_push writes the caller's frames into every rank's buffer with one
 * kernel (remote stores over NVLink).  With `active` (cvp_active_atoms, device int32) only those
 * atoms travel and land at their place in the dense buffers.  All calls are ordered on `stream`;
 * cvp_peer_signal + cvp_peer_wait form a stream-ordered barrier (every rank calls both, the same
 * number of times): after the wait, the frames pushed by all ranks before their signal are
 * visible in the local buffer.  The wait gives up after ~4 s (cvp_peer_status reports it).
 * One process per GPU; the current device at cvp_peer_buffer_create owns the buffer.             */
typedef struct cvp_peer_buffer cvp_peer_buffer;
int cvp_peer_buffer_create(int64_t bytes, cvp_peer_buffer** out);
int cvp_peer_buffer_handle(cvp_peer_buffer* buffer, void* handle64);
int cvp_peer_buffer_open(cvp_peer_buffer* buffer, int world, int rank, const void* handles /* [world][64] */);
int cvp_peer_buffer_local(cvp_peer_buffer* buffer, void** ptr, int64_t* bytes);
int cvp_peer_push(cvp_peer_buffer* buffer, int dtype, const void* src /* [count][N][3] device */,
                  int64_t first_frame, int64_t num_frames, int64_t num_atoms,
                  const int32_t* active /* device, or NULL = dense */, int32_t num_active, void* stream);
/* The same for one contiguous run of atoms [first_atom, first_atom + atom_count) of every frame:
 * 16-byte vectors when the run is aligned -- the fast form of the compact push when the active
 * atoms form a few runs (the binding splits the list).                                              */
int cvp_peer_push_rows(cvp_peer_buffer* buffer, int dtype, const void* src, int64_t first_frame,
                       int64_t num_frames, int64_t num_atoms, int64_t first_atom, int64_t atom_count,
                       void* stream);
int cvp_peer_signal(cvp_peer_buffer* buffer, void* stream);
int cvp_peer_wait(cvp_peer_buffer* buffer, void* stream);
int cvp_peer_status(cvp_peer_buffer* buffer, int* timed_out);
int cvp_peer_buffer_destroy(cvp_peer_buffer* buffer);

/* ---- trajectory ingestion (host side) --------------------------------- */
/* The callers of the reference evaluate one frame per reported step (reporting/cv_writer.py:102-109);
 * the batched analogue reads stored frames.  A DCD frame holds X, Y, Z as separate float32 records
 * in Angstrom: cvp_dcd_read maps the file and converts frames first .. first+count-1 to interleaved
 * [count][N][3] float nm (x / 10.0f) in a caller-owned host buffer (pinned memory for the
 * pipeline), split over `threads` host threads (<= 0: one per core); `cell` (may be NULL) receives
 * the raw unit-cell records [count][6] (A, gamma, B, beta, alpha, C).  No GPU involved.            */
typedef struct cvp_dcd cvp_dcd;
int cvp_dcd_open(const char* path, cvp_dcd** out);
int cvp_dcd_info(const cvp_dcd* file, int64_t* num_frames, int64_t* num_atoms, int* has_box,
                 int32_t* first_step, int32_t* interval);
int cvp_dcd_read(const cvp_dcd* file, int64_t first, int64_t count, float* positions, double* cell,
                 int threads);
int cvp_dcd_close(cvp_dcd* file);

/* GROMACS XTC (libxdrf / xdrfile compressed coordinates, big-endian XDR; nm): frames have variable
 * size, cvp_xtc_open walks the headers once and keeps an offset index; cvp_xtc_read decodes frames
 * first .. first+count-1 on `threads` host threads straight into positions [count][N][3] float nm
 * and fills box [count][9] (rows a, b, c; nm), steps [count], times [count] (ps) when not NULL.
 * cvp_xtc_write is the matching encoder (precision <= 0: 1000 per nm; box/steps/times may be NULL). */
typedef struct cvp_xtc cvp_xtc;
int cvp_xtc_open(const char* path, cvp_xtc** out);
int cvp_xtc_info(const cvp_xtc* file, int64_t* num_frames, int64_t* num_atoms, float* precision);
int cvp_xtc_read(const cvp_xtc* file, int64_t first, int64_t count, float* positions, double* box,
                 int32_t* steps, float* times, int threads);
int cvp_xtc_close(cvp_xtc* file);
int cvp_xtc_write(const char* path, int append, int64_t count, int64_t num_atoms, const float* positions,
                  const double* box, const int32_t* steps, const float* times, float precision);

/* ---- diagnostics ------------------------------------------------------------------------------- */
const char* cvp_last_error(void);
int cvp_version(void);
/* SM count and compute capability of a device; CVP_ERR_NO_DEVICE if there is none.               */
int cvp_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
/* Kernels launched by this library since the process started (all threads).                     */
int64_t cvp_launch_count(void);
/* The chunks cvp_evaluate_host cuts a batch into (frames per chunk, in order), for callers that
 * want to size pinned buffers or progress bars: `sizes` may be NULL to query `count` only.       */
int cvp_host_chunks(int64_t num_frames, int64_t frame_bytes, int64_t* sizes, int32_t capacity,
                    int32_t* count);
/* Bytes of per-spec device workspace a spec may use per evaluate call (default 4 GiB).           */
int cvp_set_workspace_limit(cvp_cv* cv, int64_t bytes);
/* Per-family tuning switch (meaning documented in DESIGN.md); returns previous value.            */
int cvp_set_option(cvp_cv* cv, const char* name, int value);
/* Named statistic of a spec.  "<kernel>_ms" -> out[0] = device milliseconds spent in that kernel
 * since the last query (CUDA events on the launching stream; needs option "time_kernels" = 1),
 * out[1] = number of launches.  Kernel names: "sweep" (pair sums), "rmsd_sums", "rmsd_gradient",
 * "rmsd_fused", "rmsd_blocks", "rmsd_multi_cov", "rmsd_multi_gradient", "gyration_sums",
 * "gyration_gradient", "gyration_fused".  Pair sums also answer "last_path": out[0] = 0 (all pairs),
 * 1 (cell-sorted, two sweeps) or 2 (cell-sorted, one symmetric sweep) for the latest evaluation.  */
int cvp_get_stat(cvp_cv* cv, const char* name, double* out);
/* Sustained FP32 FMA throughput of `device` in TFLOP/s measured by a register-resident FFMA loop
 * (the FP32 roofline denominator of the pair kernels).                                           */
int cvp_measure_fp32_peak(int device, double* tflops);
/* Same for the FP64 pipe (DFMA): the roofline denominator of the double-precision contractions
 * behind PathInRMSDSpace.                                                                        */
int cvp_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CVPACK_B200_H */
