/*
 * omega2d_b200.h -- C ABI of the B200-native Omega2D Biot-Savart direct-sum library.
 *
 * This is the drop-in boundary for ONE path of Omega2D: the O(N^2) influence sums that
 * Convection::find_vels / find_vort and solve_bem run every Runge-Kutta substep
 * (reference src/Convection.h:134-223, src/BEMHelper.h:96-99).  Each entry point below
 * replaces one reference template loop nest; citations are reference file:line.
 *
 * Conventions shared by every call (they are the reference's, src/Influence.h:82-98 and
 * src/Convection.h:202-216):
 *   - float32 structure-of-arrays, counts as int, node indices as uint32 (reference
 *     src/Omega2D.h:26 "using Int = uint32_t");
 *   - results are ACCUMULATED into the output arrays ("tu[i] += sum") on the raw,
 *     un-normalised sums: no 1/(2 pi), no freestream -- those are applied once per target
 *     collection by o2d_finalize_vels, as ElementBase::finalize_vels does;
 *   - pair terms are evaluated in float32; running sums are folded into float64 once per
 *     source tile, then added to the float32 output ("float += double"), reproducing the
 *     USE_VC-off STORE=float / ACCUM=double configuration (reference src/Simulation.h:44-50);
 *   - arrays are caller-owned and only read/written during the call; the library keeps no
 *     pointer after returning;
 *   - `mem` says where ALL array arguments of that call live: O2D_MEM_HOST (pageable or
 *     pinned host memory; the library stages H2D/D2H itself) or O2D_MEM_DEVICE (device
 *     memory of the context's GPU; the call is asynchronous on the context's stream);
 *   - every function returns O2D_OK (0) or a negative o2d_status; o2d_last_error() gives
 *     the message.  Nothing throws across this boundary.  There is NO CPU fallback: if no
 *     sm_100 device is usable, o2d_create fails.
 *   - one context per (host thread, GPU); calls on one context must be sequential, which is
 *     how the reference calls the path (src/Simulation.cpp:679-682).
 */
#ifndef OMEGA2D_B200_H
#define OMEGA2D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct o2d_context o2d_context;

typedef enum o2d_status {
  O2D_OK = 0,
  O2D_ERR_INVALID = -1,     /* bad argument (null pointer, negative count, unknown enum)   */
  O2D_ERR_UNSUPPORTED = -2, /* results type the reference asserts against (psi)            */
  O2D_ERR_CUDA = -3,        /* CUDA runtime error, message in o2d_last_error               */
  O2D_ERR_NO_DEVICE = -4,   /* no sm_100 GPU: there is no CPU fallback                     */
  O2D_ERR_COMM = -5         /* NCCL error or communicator not initialised                  */
} o2d_status;

/* core (regularisation) function: the five compile-time choices of reference
 * src/CoreFunc.h:35-39, here a run-time argument.  O2D_CORE_V2 is the reference's default. */
typedef enum o2d_core {
  O2D_CORE_RM = 0,  /* Rosenhead-Moore            src/CoreFunc.h:42-82   */
  O2D_CORE_EXP = 1, /* compact exponential        src/CoreFunc.h:85-181  */
  O2D_CORE_WL = 2,  /* Winckelmans-Leonard        src/CoreFunc.h:184-230 */
  O2D_CORE_V2 = 3,  /* Vatistas n=2 (default)     src/CoreFunc.h:233-275 */
  O2D_CORE_V3 = 4   /* Vatistas n=3               src/CoreFunc.h:278-322 */
} o2d_core;

/* results type: numeric values are the reference's results_t (src/ResultsType.h:25-31) */
typedef enum o2d_results {
  O2D_VELONLY = 1,
  O2D_VELANDGRAD = 2, /* kernelug_0v_0b, defined but never driven by the reference (Kernels.h:112-133) */
  O2D_PSIONLY = 3,    /* unsupported here exactly as in the reference (Influence.h:60)     */
  O2D_VELANDVORT = 4,
  O2D_VELANDSHEAR = 5
} o2d_results;

typedef enum o2d_mem { O2D_MEM_HOST = 0, O2D_MEM_DEVICE = 1 } o2d_mem;

/* ---- context --------------------------------------------------------------------------- */

/* Create a context on CUDA device `device` (ordinal).  Fails with O2D_ERR_NO_DEVICE when
 * the device is missing or is not compute capability 10.x. */
int o2d_create(o2d_context** ctx, int device);
int o2d_destroy(o2d_context* ctx);
/* Message of the most recent failure on this thread (never NULL). */
const char* o2d_last_error(void);
/* Use the caller's CUDA stream (a cudaStream_t passed as void*) for all work of this
 * context; NULL restores the context's own non-blocking stream. */
int o2d_set_stream(o2d_context* ctx, void* cuda_stream);
/* Block until all work queued by this context has finished. */
int o2d_synchronize(o2d_context* ctx);

/* ---- the influence sums ---------------------------------------------------------------- */

/*
 * Vortex particles on points.  Replaces points_affect_points<S,A> (reference
 * src/Influence.h:56-416; pair kernels src/Kernels.h:81-231; cores src/CoreFunc.h).
 *   sources: ns particles  sx,sy (position) ss (strength) sr (core radius)
 *   targets: nt points     tx,ty and tr (core radius) or tr==NULL for inert field points
 *            (the reference's targ.is_inert() branch, the *_0p kernels)
 *   outputs, all accumulated into:
 *     tu,tv             always
 *     tq                vorticity (O2D_VELANDVORT) or shear rate (O2D_VELANDSHEAR); else NULL
 *     tux,tvx,tuy,tvy   du/dx, dv/dx, du/dy, dv/dy (O2D_VELANDGRAD); else NULL
 * Self pairs are included exactly like the reference: they add 0 to tu,tv and the core's
 * d=0 value to tq.  With O2D_CORE_EXP the gradient-type results are NaN for coincident
 * pairs, as in the reference (src/CoreFunc.h:156-166).
 */
int o2d_points_on_points(o2d_context* ctx, int core, int results, int mem,
                         int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                         int nt, const float* tx, const float* ty, const float* tr,
                         float* tu, float* tv, float* tq,
                         float* tux, float* tvx, float* tuy, float* tvy);

/*
 * Vortex(+source) panels on points.  Replaces panels_affect_points<S,A> (reference
 * src/Influence.h:422-614; kernelu_1vs_0p src/Kernels.h:362-405, kernelu_1v_0p :312-354).
 *   panels: np panels over nnodes nodes nx,ny; idx[2j],idx[2j+1] are panel j's end nodes;
 *           vs vortex-sheet strength, ss source-sheet strength or NULL (vortex only).
 */
int o2d_panels_on_points(o2d_context* ctx, int mem,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                         const float* vs, const float* ss,
                         int nt, const float* tx, const float* ty, float* tu, float* tv);

/*
 * Vortex particles on panel centres by reciprocity.  Replaces points_affect_panels<S,A>
 * (reference src/Influence.h:633-771): pu[i] -= (1/area[i]) * sum_j kernelu_1v_0p(panel i
 * as source with strength ss[j], particle j as target).  area = Surfaces::get_area().
 */
int o2d_points_on_panels(o2d_context* ctx, int mem,
                         int ns, const float* sx, const float* sy, const float* ss,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                         const float* area, float* pu, float* pv);

/*
 * Gaussian vorticity of particles at points.  Replaces points_affect_points_vorticity<S,A>
 * (reference src/InfluenceVort.h:47-126): tw[i] += sum_j 2 ss_j exp(-d^2/sr_j^2)/sr_j^2.
 */
int o2d_points_vorticity(o2d_context* ctx, int mem,
                         int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                         int nt, const float* tx, const float* ty, float* tw);

/* ElementBase::zero_vels (reference src/ElementBase.h:200-212): u=v=0, and w,e if non-NULL. */
int o2d_zero_vels(o2d_context* ctx, int mem, int n, float* u, float* v, float* w, float* e);
/* ElementBase::finalize_vels (reference src/ElementBase.h:215-234): u = fs + u/(2 pi) in
 * double, stored float; w,e (if non-NULL) scaled by 1/(2 pi). */
int o2d_finalize_vels(o2d_context* ctx, int mem, int n, float* u, float* v, float* w, float* e,
                      double fsx, double fsy);

/* ---- device-resident evaluation and time step (one upload, one download) ------------------ */

/*
 * Convection::find_vels (reference src/Convection.h:172-223) for one particle collection and
 * one surface collection acting on one target collection:
 *     zero_vels -> points_affect_points -> panels_affect_points -> finalize_vels(fs)
 * i.e. tu,tv (and tq) are OVERWRITTEN with the finished velocities  u = fs + sum/(2 pi)
 * (vorticity / shear scaled by 1/(2 pi)).  np == 0 means no surface collection; pss == NULL
 * means vortex-only panels; tr == NULL means inert targets.  results: O2D_VELONLY,
 * O2D_VELANDVORT or O2D_VELANDSHEAR (panels contribute to tu,tv only, as in the reference,
 * src/Influence.h:422-614).  With O2D_MEM_HOST the inputs are uploaded once and only the
 * results come back; the intermediate zero/accumulate/finalize passes never leave the GPU.
 */
int o2d_find_vels(o2d_context* ctx, int core, int results, int mem,
                  int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                  int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                  const float* pvs, const float* pss,
                  int nt, const float* tx, const float* ty, const float* tr,
                  float* tu, float* tv, float* tq, double fsx, double fsy);

/*
 * Convection::find_vels for ANY number of source collections: the reference loops "for src in vort: visit;
 * for src in bdry: visit" (src/Convection.h:205-213), i.e. one influence call per (source, target) pair.  Here every
 * Points source is packed behind the previous one and summed by ONE pair-sum launch, every Surfaces source likewise by
 * ONE panel launch, around one zero and one finalize: with host arrays that is one upload of all collections and one
 * download of the target's results however many collections the simulation holds (particles of several bodies, Points
 * boundaries, several Surfaces).  Otherwise as o2d_find_vels.
 */
typedef struct o2d_points_view { int n; const float* x; const float* y; const float* s; const float* r; } o2d_points_view;
typedef struct o2d_panels_view {
  int np; int nnodes; const float* nx; const float* ny; const uint32_t* idx; const float* vs; const float* ss;
} o2d_panels_view;
int o2d_find_vels_multi(o2d_context* ctx, int core, int results, int mem,
                        int nsrc, const o2d_points_view* src, int npan, const o2d_panels_view* pan,
                        int nt, const float* tx, const float* ty, const float* tr,
                        float* tu, float* tv, float* tq, double fsx, double fsy);

/*
 * One convection step of a Lagrangian particle collection, all stages on the GPU:
 * Convection::advect_1st (order 1, reference src/Convection.h:277-308), advect_2nd_ralston
 * (order 2, :314-392) or advect_3rd (order 3, :472-595) with ElementBase::move
 * (src/ElementBase.h:276-359) between the find_vels evaluations.  x,y are advanced in place;
 * u,v receive what the reference leaves in the collection's velocity arrays (the weighted
 * stage velocities for order 2 and 3).  For particle-only flows (np == 0, e.g. the reference's
 * examples/speedtest20kx10.json) this is the reference's whole convection step.  With np > 0 the
 * surface strengths pvs/pss are held FIXED over the step and clear_inner_layer is NOT run -- a
 * frozen-body approximation kept for callers that handle both themselves; the reference's own
 * step with bodies (BEM re-solved and inner layer cleared after every stage) is o2d_advect_bodies.
 */
int o2d_advect(o2d_context* ctx, int core, int order, int mem,
               int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
               int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
               const float* pvs, const float* pss, double fsx, double fsy, double dt);

/*
 * `nsteps` consecutive o2d_advect steps with the particle state resident on the GPU in between: what
 * Simulation::step does nsteps times for an inviscid particle-only run (reference src/Simulation.cpp:687-744 with
 * "viscous": "none", e.g. examples/speedtest20kx10.json = 10 steps of 20 000 particles), with one upload before
 * the first and one download after the last step.  u,v hold what the last step leaves in them.
 */
int o2d_advect_steps(o2d_context* ctx, int core, int order, int mem,
                     int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                     int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                     const float* pvs, const float* pss, double fsx, double fsy, double dt, int nsteps);

/*
 * One convection step of a flow WITH bodies, particles resident on the GPU across all Runge-Kutta stages.
 *
 * The reference's stage is find_derivs = solve_bem -> find_vels (src/Convection.h:229-243) followed by
 * ElementBase::move and clear_inner_layer (:303-307, :334, :388-391, :489-578).  solve_bem's linear solve belongs
 * to the host application (Eigen GMRES, src/BEM.h:166-232); everything around it is done here:
 *   per stage   zero_vels + points_affect_panels + Surfaces::finalize_vels(fs) on the panels (src/BEMHelper.h:88-100)
 *               -> `solve` (host callback: right-hand side, linear solve, new sheet strengths)
 *               -> find_vels of the particles on themselves with the new strengths -> move -> clear_inner_layer(1, ...)
 * so a step costs ONE upload and ONE download of the particle arrays plus, per stage, 2*Np floats each way.
 *
 *   solve(user, stage, time, np, pu, pv, vs, ss): pu,pv = finalised particle-induced velocity at the panel centres
 *       (what Surfaces::get_vel holds when the reference calls vels_to_rhs_panels); writes the new vortex strengths to
 *       vs[np] and -- when the caller passed pss != NULL -- the source strengths to ss[np]; returns 0 on success.
 *       vs/ss are the caller's own pvs/pss arrays, so after the call they hold the strengths of the last stage.
 *       stage = 0, 1, 2; time = the stage time the reference passes to solve_bem.
 *   area, snx, sny: panel lengths and unit normals (Surfaces::get_area / get_norm); cutoff_mult = 0.5/sqrt(2 pi) and
 *       ips as for o2d_clear_inner_layer.  Bodies do not move inside a step (fixed or externally transformed panels).
 *   nf, fx, fy, fu, fv: an optional collection of inert Lagrangian field points (tracers; nf = 0: none).  The reference
 *       evaluates and moves them alongside the particles in every stage ("find_vels(_fs, _vort, _bdry, _fldpt)",
 *       src/Convection.h:242, the interim_fldpt copies :343-348) and clears them out of the body too; here they ride
 *       along on the device: positions advance in place, fu,fv receive what the reference leaves in their velocity arrays.
 * Host arrays only (the callback runs on the host).  x,y are advanced in place; u,v as for o2d_advect.
 */
typedef int (*o2d_bem_solve_fn)(void* user, int stage, double time, int np, const float* pu, const float* pv,
                                float* vs, float* ss);
int o2d_advect_bodies(o2d_context* ctx, int core, int order, int mem,
                      int n, float* x, float* y, const float* s, const float* r, float* u, float* v,
                      int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                      const float* area, const float* snx, const float* sny, float* pvs, float* pss,
                      double fsx, double fsy, double time, double dt, float cutoff_mult, float ips,
                      int nf, float* fx, float* fy, float* fu, float* fv,
                      o2d_bem_solve_fn solve, void* user);

/*
 * Surfaces::add_solved_rot_strengths / add_rot_strengths_base (reference src/Surfaces.h:556-620), which
 * find_vels applies with factor +1 before and -1 after the sums (src/Convection.h:186-189,219-222):
 * adds to vs,ss the vortex and source sheet strengths equivalent to the solid-body rotation of the
 * panels' body at angular rate `rotvel` (= factor * rotation rate) about its geometric centre (cx,cy),
 * from the body's UNTRANSFORMED node coordinates ux,uy and the panel lengths `area`.
 */
int o2d_add_rot_strengths(o2d_context* ctx, int mem, int np, int nnodes, const float* ux, const float* uy,
                          const uint32_t* idx, const float* area, float cx, float cy, float rotvel,
                          float* vs, float* ss);

/* ---- BEM influence coefficients ------------------------------------------------------------ */

/*
 * Panel-on-panel influence-coefficient block.  Replaces the O(Np^2) loop nest of
 * panels_on_panels_coeff<S> (reference src/Coefficients.h:155-427; pair kernels kernelu_1vos_0p
 * src/Kernels.h:413-457 and kernelu_1v_0p :312-354), which solve_bem calls through
 * CoefficientVisitor for every (source, target) pair of surfaces whenever the geometry changes
 * (src/BEMHelper.h:165-226).
 *   source surface: nsrc panels over nsnodes nodes (snx,sny), end nodes sidx[2j],sidx[2j+1],
 *                   panel lengths sarea (Surfaces::get_area)
 *   target surface: same, plus unit tangents ttx,tty (Surfaces::get_tang) and normals tnx,tny
 *                   (Surfaces::get_norm; may be NULL when both_strengths == 0)
 *   both_strengths: 1 = vortex AND source sheet strengths unknown on both surfaces (what the
 *                   reference always runs: 2 columns per source panel, 2 rows per target panel);
 *                   0 = vortex strengths only (1 column / 1 row per panel)
 *   same_surface:   1 when source and target are the same collection (self influence i == j is
 *                   set to 1/2 on the diagonal entries, src/Coefficients.h:392-404)
 *   coeffs:         column-major, leading dimension ld >= rows (rows = (both ? 2 : 1) * ntarg):
 *                   coeffs[col * ld + row], col = 2j (+1 for the source strength), row = 2i (+1 for
 *                   the normal condition).  Already scaled by 1/(2 pi), like the reference's return
 *                   value.  Only the rows x cols block is written, so ld = rows + 1 lets the caller
 *                   append the reference's augmentation row/column (src/Coefficients.h:431-530)
 *                   itself -- that part is O(Np) and stays on the host.  With O2D_MEM_DEVICE and
 *                   both_strengths, ld must be even and coeffs 8-byte aligned.
 */
int o2d_panels_on_panels_coeff(o2d_context* ctx, int mem,
                               int nsrc, int nsnodes, const float* snx, const float* sny,
                               const uint32_t* sidx, const float* sarea,
                               int ntarg, int ntnodes, const float* tnx_nodes, const float* tny_nodes,
                               const uint32_t* tidx, const float* tarea,
                               const float* ttx, const float* tty, const float* tnx, const float* tny,
                               int both_strengths, int same_surface, long long ld, float* coeffs);

/* ---- clear_inner_layer ------------------------------------------------------------------ */

/*
 * Push particles out of the innermost layer around a body.  Replaces clear_inner_panp2<S> with
 * _method == 1 (reference src/Reflect.h:490-733; closest-point kernel panel_point_distance
 * :108-170) as called by clear_inner_layer<S> (:735-768) for one Surfaces boundary and one
 * Lagrangian Points collection -- what Convection does after every move of every Runge-Kutta stage
 * (src/Convection.h:303-307) with cutoff_mult = 0.5/sqrt(2 pi) and ips = the inter-particle spacing.
 *   panels: np panels over nnodes nodes (nx,ny), idx, unit panel normals snx,sny (Surfaces::get_norm)
 *   points: n positions tx,ty, moved in place (strengths and radii are not touched by method 1)
 *   num_moved: if non-NULL receives the number of pushes (the reference's "pushed N particles")
 * Every particle closer than cutoff_mult*ips (minus 0.001 ips) to the surface along the mean normal of
 * its closest panel(s)/node(s) is moved onto that layer; the search repeats for moved particles until
 * none moves (corners).  method 0 (strength cropping with the cut tables, src/Reflect.h:392-480,
 * :668-690) has no caller anywhere in the reference and returns O2D_ERR_UNSUPPORTED.
 */
int o2d_clear_inner_layer(o2d_context* ctx, int method, int mem,
                          int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                          const float* snx, const float* sny,
                          int n, float* tx, float* ty, float cutoff_mult, float ips, long long* num_moved);

/*
 * Mirror particles that ended up inside a body back to the outside.  Replaces reflect_panp2<S>
 * (reference src/Reflect.h:174-350) as called by reflect_interior<S> (:357-380) for one Surfaces
 * boundary and one Points collection -- what Diffusion::step does before VRM, which "only works in
 * free space" (src/Diffusion.h:299-301).  Same panel and point arguments as o2d_clear_inner_layer.
 * For each particle: closest panel(s)/node(s) (exact ties averaged), mean normal and mean contact
 * point cp; if (x - cp).normal < 0 the particle is moved to cp + dist*normal.  One pass (no re-test).
 *   num_reflected: if non-NULL receives the reference's "reflected N particles" count.
 */
int o2d_reflect_interior(o2d_context* ctx, int mem,
                         int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                         const float* snx, const float* sny,
                         int n, float* tx, float* ty, long long* num_reflected);

/* ---- status reductions ----------------------------------------------------------------------- */

/*
 * The per-collection sums Simulation::write_status and Points::update_max_str need (reference
 * src/Simulation.cpp:765-805): total circulation sum s_i (ElementBase::get_total_circ,
 * src/ElementBase.h:374-394), impulse (-sum s_i y_i, +sum s_i x_i) (Points::get_total_impulse,
 * src/Points.h:270-285) and max |s_i| (ElementBase::get_max_str, :362-372), in one pass over
 * device-resident (or host) arrays.  Sums are float64 in a fixed order, returned as float like the
 * reference's S; any output pointer may be NULL.  Synchronous: it returns four scalars.
 */
int o2d_points_status(o2d_context* ctx, int mem, int n, const float* x, const float* y, const float* s,
                      float* total_circ, float* impulse_x, float* impulse_y, float* max_str);

/* ---- .vtu output of a particle collection ----------------------------------------------------- */

/*
 * The .vtu file image Points<S>::write_vtk produces (reference src/Points.h:556-740; VtkXmlWriter::unpackArray /
 * writeDataArray src/VtkXmlWriter.h:71-106, base64 through reference lib/cppcodec/base64_rfc4648.hpp), byte for
 * byte: XML skeleton as tinyxml2 prints it, every DataArray as ' ' + base64(UInt32 length of the encoded text) +
 * base64(data) + ' '.  The arrays are encoded on the GPU straight from the SoA arrays (the (x,y,0) interleave and
 * the connectivity/offsets/types arrays never exist in memory); only the finished text crosses PCIe.
 *   x,y,u,v  positions and velocities (n each; `mem` says where they live)
 *   s,r      circulation and radius, both NULL for inert field points (src/Points.h:561-568)
 *   w,e      vorticity / shear rate or NULL (ElementBase::has_vort / has_shear, src/ElementBase.h:63,96)
 *   time     the FieldData TimeValue
 *   out,cap  HOST buffer receiving the image (always host memory: it is a file); *written = its size, also
 *            set when cap is too small (then O2D_ERR_INVALID is returned and nothing is written)
 * o2d_vtu_points_size gives the exact size beforehand.  n must be > 0 (the reference asserts it) and < 2^28
 * (the format's UInt32 array header).  o2d_write_vtu_points writes the image to `path` through a pinned
 * buffer owned by the context; the caller composes the reference's file name ("part_%02d_%05d.vtu" /
 * "fldpt_...", src/Points.h:570-573).
 */
int o2d_vtu_points_size(int n, int has_s, int has_r, int has_w, int has_e, size_t* bytes);
int o2d_vtu_points(o2d_context* ctx, int mem, int n, const float* x, const float* y,
                   const float* s, const float* r, const float* w, const float* e,
                   const float* u, const float* v, double time,
                   char* out, size_t cap, size_t* written);
int o2d_write_vtu_points(o2d_context* ctx, int mem, const char* path, int n, const float* x, const float* y,
                         const float* s, const float* r, const float* w, const float* e,
                         const float* u, const float* v, double time);

/* ---- the reference's existing plugin seam ---------------------------------------------- */

/*
 * Drop-in definitions of the two symbols Omega2D binds when built with
 * -DEXTERNAL_VEL_SOLVE (reference src/Influence.h:31-35, call sites :82-88; the stock
 * implementation is onbody's treecode, extern/onbody/src/interface2dvort.cpp:182-187 and
 * interface2dvorttr.cpp:177-183).  Same Fortran-style signature (counts by pointer, note
 * ss BEFORE sr), same semantics: velocities are accumulated into tu,tv, the return value is
 * the flop count the caller turns into its GFlop/s print (negative on failure).  They use a
 * lazily created process-wide context on device $O2D_DEVICE (default 0) -- or, with $O2D_DEVICES=N|all, a
 * process-wide o2d_group over N (all visible) GPUs -- and the core set by o2d_set_default_core / $O2D_CORE
 * (rm|exp|wl|v2|v3; default v2 like the reference).
 */
float external_vel_solver_f_(int* ns, const float* sx, const float* sy, const float* ss, const float* sr,
                             int* nt, const float* tx, const float* ty, float* tu, float* tv);
float external_vel_solver_tr_f_(int* ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                int* nt, const float* tx, const float* ty, const float* tr,
                                float* tu, float* tv);
int o2d_set_default_core(int core);
int o2d_get_default_core(void);

/* ---- multi-GPU: targets sharded in contiguous blocks, sources all-gathered -------------- */

/* 128-byte NCCL unique id; rank 0 creates it, the host program distributes it. */
#define O2D_UNIQUE_ID_BYTES 128
int o2d_comm_unique_id(void* id128);
/* Join communicator of `nranks` ranks as `rank` (one context / process / GPU per rank). */
int o2d_comm_init(o2d_context* ctx, int nranks, int rank, const void* id128);
int o2d_comm_destroy(o2d_context* ctx);
/*
 * Sharded points-on-points: this rank owns ns_local source particles (its block of the
 * global set) and nt target points.  Each rank's packed source block is replicated to every
 * rank with ONE in-place ncclAllGather over NVLink (16 B per particle) on the launching stream,
 * and the gathered array is summed by the same single launch as on one GPU (the gather is
 * ~1e-4 of the evaluation; a launch that owns the whole device is worth more than hiding it);
 * outputs stay on the owning rank.  ns_block is the per-rank block
 * capacity agreed by all ranks (>= every rank's ns_local, e.g. ceil(N/nranks)).
 * Same argument meaning as o2d_points_on_points otherwise.
 */
int o2d_points_on_points_sharded(o2d_context* ctx, int core, int results, int mem,
                                 int ns_local, int ns_block,
                                 const float* sx, const float* sy, const float* ss, const float* sr,
                                 int nt, const float* tx, const float* ty, const float* tr,
                                 float* tu, float* tv, float* tq,
                                 float* tux, float* tvx, float* tuy, float* tvy);

/*
 * Sharded Convection::find_vels: o2d_find_vels with the particle collection split over the ranks
 * like o2d_points_on_points_sharded (this rank's block of nv_local <= nv_block particles is
 * all-gathered once), the surface collection replicated on every rank (it is small: 24 B per
 * panel) and this rank's nt targets.  tu,tv,tq are overwritten with the finished velocities.
 */
int o2d_find_vels_sharded(o2d_context* ctx, int core, int results, int mem,
                          int nv_local, int nv_block,
                          const float* vx, const float* vy, const float* vs, const float* vr,
                          int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                          const float* pvs, const float* pss,
                          int nt, const float* tx, const float* ty, const float* tr,
                          float* tu, float* tv, float* tq, double fsx, double fsy);
/*
 * Sharded points_affect_panels (the BEM right-hand side, reference src/BEMHelper.h:96-99): every
 * rank holds ns_local particles of the global set and ALL np panels; the per-panel reciprocity
 * sums are added over the ranks with one ncclAllReduce of 2*np float64 and folded into pu,pv
 * ("pu -= sum/area") identically on every rank.
 */
int o2d_points_on_panels_sharded(o2d_context* ctx, int mem,
                                 int ns_local, const float* sx, const float* sy, const float* ss,
                                 int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                 const float* area, float* pu, float* pv);

/* ---- all GPUs of the box from ONE process ------------------------------------------------------- */

/*
 * Omega2D is a single-process OpenMP program whose influence loops are "omp parallel for" over the targets
 * (reference src/Influence.h:198,348).  A group applies the same decomposition to devices: HOST arrays in, the
 * target collection cut into one contiguous block per GPU (multiples of 1024), the whole source collection
 * uploaded to every GPU over its own PCIe link, one host thread per device driving the single-device entry
 * point above on its block, results written to disjoint slices of the caller's arrays.  Argument meaning and
 * accumulate / overwrite semantics are those of the single-device call of the same name.  Small problems
 * (fewer than `pairs_per_device` pair evaluations per device, default 2e9; a panel pair counts as five) use
 * fewer devices, down to one.  o2d_group_points_on_panels cuts the SOURCES instead (targets are the few
 * panels) and adds the devices' per-panel sums on the host in fixed order.
 *   devices == NULL: devices 0..ndev-1;  ndev <= 0: every visible device.
 */
typedef struct o2d_group o2d_group;
int o2d_group_create(o2d_group** group, int ndev, const int* devices);
int o2d_group_destroy(o2d_group* group);
int o2d_group_size(const o2d_group* group);
o2d_context* o2d_group_context(o2d_group* group, int slot);   /* the slot's context, e.g. for the other calls */
int o2d_group_set_min_pairs(o2d_group* group, long long pairs_per_device);
int o2d_group_points_on_points(o2d_group* group, int core, int results,
                               int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                               int nt, const float* tx, const float* ty, const float* tr,
                               float* tu, float* tv, float* tq,
                               float* tux, float* tvx, float* tuy, float* tvy);
int o2d_group_find_vels(o2d_group* group, int core, int results,
                        int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                        int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                        const float* pvs, const float* pss,
                        int nt, const float* tx, const float* ty, const float* tr,
                        float* tu, float* tv, float* tq, double fsx, double fsy);
int o2d_group_panels_on_points(o2d_group* group, int np, int nnodes, const float* nx, const float* ny,
                               const uint32_t* idx, const float* vs, const float* ss,
                               int nt, const float* tx, const float* ty, float* tu, float* tv);
int o2d_group_points_on_panels(o2d_group* group, int ns, const float* sx, const float* sy, const float* ss,
                               int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                               const float* area, float* pu, float* pv);

/* ---- measurement hooks (used by bench.py; no effect on results) ------------------------- */

/* Reference ("algorithmic") flops per pair, from the reference's own counters
 * (src/Kernels.h flops*_*(), src/CoreFunc.h flops_t{p,v}_{no,}grads()).
 * kind: 0 points-on-points (uses core, results, blob_targets), 1 panels vortex-only,
 * 2 panels vortex+source, 3 Gaussian vorticity. */
int o2d_flops_per_pair(int kind, int core, int results, int blob_targets);
/* Kernel launch counter and CUDA-event timing of the dominant (pair-sum) kernels, measured
 * on the stream they are launched on.  Timing is off by default. */
int o2d_enable_timing(o2d_context* ctx, int on);
int o2d_reset_counters(o2d_context* ctx);
int o2d_get_counters(o2d_context* ctx, long long* kernel_launches, long long* pair_kernel_launches,
                     double* pair_kernel_ms);
/* Device facts: SM count, max SM clock (MHz), nominal FP32 peak = SMs*128*2*clock (TFLOP/s). */
int o2d_device_info(o2d_context* ctx, int* sm_count, int* sm_clock_mhz, double* fp32_peak_tflops);
/* Sustained FP32 throughput of a dependent-chain-free FFMA (packed=0) or FFMA2 (packed=1)
 * micro-kernel on this device, in TFLOP/s (2 flops per lane-FMA). */
int o2d_measure_fp32_peak(o2d_context* ctx, int packed, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* OMEGA2D_B200_H */
