/* lgs_oracle.h -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference hot path.
 *
 * Every function restates one piece of /root/reference (file:line cited at each definition in
 * lgs_oracle.c) on dense row-major double grids (0.0 = unknown, out-of-map reads = 0.0, which
 * is what GridMap::Value(x, y, unknown) returns, grid_map.hpp:859-873).
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * port is pinned against the reference ITSELF: tests/test_oracle_port.py compares it with
 * oracle/_ref/liblgs_ref.so (the unmodified reference objects) on seeded random inputs wherever
 * that library exists, and against tests/golden/ (vectors generated from the same library by
 * tests/golden/make_golden.py) everywhere.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it; the product never does.
 */
#ifndef LGS_ORACLE_H
#define LGS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_geom {
    int nx, ny;
    double min_x, min_y, res;
    int patch;
} orc_geom;

typedef struct orc_match {
    int found, ix, iy, it;
    int win_x, win_y, win_t, pad;
    double step_x, step_y, step_t;
    double score;
    double sensor_pose[3];
    double best_sensor_pose[3];
    long long n_scored;            /* score evaluations the CPU search performed */
} orc_match;

double orc_bayes_update(double value, double prob);
int orc_bresenham(int x0, int y0, int x1, int y1, int* out_xy, int cap);
void orc_sliding_window_max(const double* in, int n, int w, double* out);
void orc_precompute(const double* grid, int nx, int ny, int w, double* out);
void orc_pyramid(const double* grid, int nx, int ny, int height_max, double* out);

void orc_compound(const double* a, const double* b, double* o);
int orc_hit_points(const double* robot_pose, const double* rel, int n, const double* angles,
                   const double* ranges, double scan_min_range, double scan_max_range,
                   double usable_min, double usable_max, double* sensor_pose, double* hit_xy,
                   double* bbox);
void orc_geometry_resize(const orc_geom* cur, double min_x, double min_y, double max_x,
                         double max_y, orc_geom* out, int* shift_x, int* shift_y);
int orc_geometry_expand(const orc_geom* cur, double min_x, double min_y, double max_x,
                        double max_y, double enlarge_step, orc_geom* out, int* shift_x,
                        int* shift_y);
int orc_integrate_hits(double* grid, const orc_geom* g, const double* sensor_xy, int n,
                       const double* hit_xy, double p_hit, double p_miss);

int orc_rtcsm_match(const double* grid, const double* coarse, const orc_geom* g, int low_res,
                    double range_x, double range_y, double range_theta, double scan_range_max,
                    const double* init_pose, const double* rel, int n, const double* angles,
                    const double* ranges, double norm_threshold, orc_match* out);
double orc_pixel_accurate_score(const double* level, const orc_geom* g, double usable_min,
                                double usable_max, const double* sensor_pose, int n,
                                const double* angles, const double* ranges,
                                double scan_min_range, double scan_max_range);
double orc_cost_greedy_endpoint(const double* grid, const orc_geom* g, const double* cost,
                                const double* sensor_pose, int n, const double* angles,
                                const double* ranges, double scan_min_range, double scan_max_range);
void orc_cost_tail(const double* grid, const orc_geom* g, const double* cost, const double* best_pose,
                   int n, const double* angles, const double* ranges, double scan_min_range,
                   double scan_max_range, double* normalized_cost, double* cov);
int orc_gs_match(const double* grid, const orc_geom* g, double range_x, double range_y, double range_theta,
                 double step_x, double step_y, double step_theta, double usable_min, double usable_max,
                 const double* init_pose, const double* rel, int n, const double* angles,
                 const double* ranges, double scan_min_range, double scan_max_range,
                 double norm_threshold, orc_match* out);
int orc_bb_match(const double* pyramid, const orc_geom* g, int height_max, double range_x,
                 double range_y, double range_theta, double scan_range_max, double usable_min,
                 double usable_max, const double* init_pose, const double* rel, int n,
                 const double* angles, const double* ranges, double scan_min_range,
                 double scan_max_range, double norm_threshold, orc_match* out);

#ifdef __cplusplus
}
#endif
#endif
