/*
 * ftgp_oracle.h -- CPU restatement (plain C, double precision) of the
 * ft_grandprix hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker (or the
 * timed CPU baseline).  Nothing under ft_grandprix_b200/ may import, link or
 * call it: the product path is the CUDA library behind include/ftgp.h.
 *
 * PARITY PIN STATUS
 *   - track compiler  : pinned against ft_grandprix.chunk.chunk() outputs
 *                       (tests/golden/chunks_*.json, made by
 *                       tests/golden/make_golden.py from /root/reference).
 *   - drivers         : pinned against ft_grandprix.nidc / ft_grandprix.fast
 *                       (tests/golden/drivers.npz).
 *   - lap logic       : restated from ft_grandprix/custom.py:1340-1372, pinned
 *                       against a line-by-line Python replay of that block in
 *                       tests (the block itself cannot be imported without
 *                       mujoco + dearpygui).
 *   - mj_ray / mj_step: "parity unpinned".  The arithmetic lives in the
 *                       third-party `mujoco` wheel (3.2.2 per requirements.txt:4,
 *                       3.3.2 per uv.lock:104-105), which is absent from
 *                       /root/reference and from this image.  These functions
 *                       restate MuJoCo's published algorithm (SURVEY.md
 *                       Appendix B) for the model in template/mushr.em.xml.
 *   - centreline      : restates svg.path 6.3 (requirements.txt:5) Path.point();
 *                       svg.path is absent here, so also unpinned.
 */
#ifndef FTGP_ORACLE_H
#define FTGP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTO_NBEAMS 90
#define FTO_NQ 34
#define FTO_NV 29
#define FTO_NPATH 100

/* ---------------------------------------------------------------- track */
typedef struct fto_track fto_track;

/* a8 + a9: ft_grandprix/chunk.py:39-64, template/mushr.em.xml:19-20,55,92.
 * wall: uint8[h*w] row-major, non-zero = wall pixel (R+G+B == 765). */
fto_track* fto_track_create(const uint8_t* wall, int w, int h, double scale, int chunk_px);
void fto_track_destroy(fto_track* t);
int fto_track_nchunks(const fto_track* t);
void fto_track_dims(const fto_track* t, int* hc, int* vc, double* size_x, double* size_y);
/* chunk list in chunk.py's scan order: out[2*k] = i, out[2*k+1] = j */
void fto_track_chunks(const fto_track* t, int32_t* out);
/* hfield elevation of chunk k after MuJoCo's flip + normalise: out[nrow*ncol] */
void fto_track_chunk_data(const fto_track* t, int k, float* out, int* nrow, int* ncol);

/* ---------------------------------------------------------------- centreline */
/* a10: ft_grandprix/curve.py:6-18 + custom.py:1184-1186 (svg.path semantics).
 * d: SVG path data string.  out: double[npoints*2], metres.  Returns 0 on ok. */
int fto_centreline(const char* d, int npoints, int img_w, int img_h, int chunk_w, int chunk_h,
                   double scale, double* out);

/* ---------------------------------------------------------------- lidar */
/* a3: mj_ray for one car's 90 rangefinder sites (template/mushr.em.xml:112-117,204-206).
 * pose: qpos[0:7] of the free joint (pos, quat w-first).  out: double[90], -1 on miss. */
void fto_lidar_scan(const fto_track* t, const double* pose, double* out);
/* generic single ray against walls + ground plane (mj_ray restated) */
double fto_ray(const fto_track* t, const double* pnt, const double* vec);
/* multi-car world: other cars' lidar-visible geoms.  poses: double[ncars*7] root poses
 * of every car in the world, self = index of the scanning car; visible: uint8[ncars]
 * (0 = shadowed, custom.py:1455-1464).  Only the lidar cylinder of other cars is
 * above the ray plane in level driving; wheels/chassis are tested too. */
void fto_lidar_scan_world(const fto_track* t, const double* qpos_all, int ncars, int self,
                          const uint8_t* visible, double* out);
/* the same with full state rows: qpos_all [ncars][stride]; stride >= 34: the other cars' wheel poses follow their
 * suspension / steering joints; the other cars' lidar cylinder, chassis mesh triangles and wheel ellipsoids are targets */
void fto_lidar_scan_world_stride(const fto_track* t, const double* qpos_all, int stride, int ncars, int self,
                                 const uint8_t* visible, double* out);

/* ---------------------------------------------------------------- drivers */
/* a5: ft_grandprix/nidc.py:116-131 (kind 0), ft_grandprix/fast.py:118-139 (kind 1),
 * lobotomy (kind 2).  Returns 0, or 1 when the Python driver would raise
 * (NaN -> ValueError in int(np.ceil(nan)); custom.py:1409-1411 keeps old ctrl). */
int fto_driver(int kind, const double* ranges, int n, double* speed, double* steer);

/* ---------------------------------------------------------------- lap logic */
typedef struct {
    int32_t offset, completion, laps, start, good_start, finished, ntimes, off_track;
    int32_t rank;           /* winners[id], 0 = not finished */
    int32_t delta;
} fto_lap_state;
/* a2: ft_grandprix/custom.py:1340-1372.  times: int32[max_times] lap durations in steps.
 * nwinners: in/out world counter (len(self.winners)). */
void fto_lap_update(fto_lap_state* s, int32_t* times, int max_times, const double* path,
                    const double* xy, int32_t steps, int32_t lap_target, int32_t* nwinners);

/* ---------------------------------------------------------------- vehicle step */
typedef struct fto_model fto_model;
/* model of template/mushr.em.xml for one car (ncars cars per world share nothing but
 * car-car contacts).  Derived constants (inertias, invweight0) computed here. */
fto_model* fto_model_create(void);
void fto_model_destroy(fto_model* m);
/* compile-time constants, for cross-checking the product's generated header */
void fto_model_constants(const fto_model* m, double* dof_invweight0 /*29*/,
                         double* body_invweight0 /*2*11*/, double* body_mass /*11*/,
                         double* body_inertia /*11*9 about CoM, body frame*/,
                         double* body_ipos /*11*3*/, double* meaninertia);
/* a7: one mj_step for one independent car.  qpos[34], qvel[29], warm[29] in/out;
 * ctrl[2] = (forward, turn) (custom.py:1422-1423).  t may be NULL (no walls).
 * info (optional, int[8]): [0]=newton iters, [1]=nefc, [2]=ncon wheel-plane,
 * [3]=ncon wall, [4]=ncon chassis/lidar-plane.  Returns 0 ok, 1 if MuJoCo would
 * have reset the data (NaN / |x|>1e10). */
int fto_step(const fto_model* m, const fto_track* t, double* qpos, double* qvel,
             double* warm, const double* ctrl, int* info);
/* mj_resetData + position_vehicles for one car (custom.py:1092,1232-1245) */
void fto_reset(const fto_model* m, double* qpos, double* qvel, double* warm,
               double x, double y, double yaw);
/* pieces, exposed for unit tests */
void fto_mass_matrix(const fto_model* m, const double* qpos, double* M /*29*29*/);
void fto_bias(const fto_model* m, const double* qpos, const double* qvel, double* bias /*29*/);
void fto_inverse(const fto_model* m, const double* qpos, const double* qvel,
                 const double* qacc, double* tau /*29*/);
double fto_energy(const fto_model* m, const double* qpos, const double* qvel);


/* TEST SUPPORT: the convex problem of mj_fwdConstraint for one state (see oracle/step.c) */
int fto_constraint_problem(const fto_model* m, const fto_track* t, const double* qpos, const double* qvel, const double* ctrl,
                           int maxrows, double* M, double* qfrc_smooth, double* J, double* D, double* R, double* aref,
                           double* floss, int* type, double* pos);


/* N-car world (BASELINE config 5): one Newton problem over all cars, car-car contacts by this framework's definition
 * (oracle/step.c).  Ground work for the coupled solve; not used by the product yet. */
/* option bubble_wrap (custom.py:970-972,1041-1055): the softener spheres collide with the walls */
void fto_model_set_bubble_wrap(fto_model* m, int on);
/* TEST SUPPORT: contact set of one car (10 doubles per contact: body, dist, pos[3], normal[3], mu, solimp d0) */
int fto_contacts(const fto_model* m, const fto_track* t, const double* qpos, double* out, int maxcon);
int fto_world_step(const fto_model* m, const fto_track* t, int ncars, double* qpos, double* qvel, double* warm,
                   const double* ctrl, const uint8_t* shadowed, int* info);
int fto_world_problem(const fto_model* m, const fto_track* t, int ncars, const double* qpos, const double* qvel, const double* ctrl,
                      int maxrows, double* M, double* qfs, double* J, double* D, double* R, double* aref, double* floss, int* type,
                      int* ncc_out);

#ifdef __cplusplus
}
#endif
#endif
