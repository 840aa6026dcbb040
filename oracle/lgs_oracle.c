/* lgs_oracle.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference hot path.
 * See lgs_oracle.h for how it is pinned.  Compile with -ffp-contract=off (no FMA), like the
 * reference objects.  All "file:line" citations are relative to /root/reference. */
#include "lgs_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- grid access ------------------------------------------------------------------------------ */

/* GridMap::Value(x, y, unknown): include/my_lidar_graph_slam/grid_map/grid_map.hpp:859-873
 * (outside the map or in an unallocated patch -> the unknown value 0.0). */
static double grid_value(const double* g, int nx, int ny, int x, int y) {
    if (x < 0 || x >= nx || y < 0 || y >= ny) return 0.0;
    return g[(size_t)y * nx + x];
}

/* GridMap::WorldCoordinateToGridCellIndex: grid_map.hpp:779-790 */
static int world_to_cell(double p, double min_p, double res) {
    return (int)floor((p - min_p) / res);
}

/* GridMap::GridCellIndexToPatchIndex: grid_map.hpp:907-916 (idx / size - 1 for idx < 0) */
static int cell_to_patch(int idx, int patch) { return idx < 0 ? idx / patch - 1 : idx / patch; }

/* ---- BinaryBayesGridCell: grid_map/binary_bayes_grid_cell.hpp ----------------------------------- */

static double clamp_prob(double v) {                       /* :95-101, constants :50-52 */
    const double lo = 1e-3, hi = 1.0 - 1e-3;
    return v < lo ? lo : (hi < v ? hi : v);
}
static double value_to_odds(double v) {                    /* :104-113 */
    const double c = clamp_prob(v);
    return c / (1.0 - c);
}
static double odds_to_value(double o) { return clamp_prob(o / (1.0 + o)); }   /* :116-119 */

double orc_bayes_update(double value, double prob) {       /* Update, :75-92 */
    if (value == 0.0) return clamp_prob(prob);
    const double old_odds = value_to_odds(value);
    const double value_odds = value_to_odds(prob);
    const double new_value = odds_to_value(old_odds * value_odds);
    return clamp_prob(new_value);
}

/* ---- Bresenham: include/my_lidar_graph_slam/util.hpp:257-303 ------------------------------------- */

int orc_bresenham(int x0, int y0, int x1, int y1, int* out, int cap) {
    int dx = x1 - x0, dy = y1 - y0;
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    int x = x0, y = y0, n = 0;
    dx = abs(dx * 2);
    dy = abs(dy * 2);
#define EMIT() do { if (n < cap) { out[2 * n] = x; out[2 * n + 1] = y; } ++n; } while (0)
    EMIT();
    if (dx > dy) {
        int err = dy - dx / 2;
        while (x != x1) {
            if (err >= 0) { y += sy; err -= dx; }
            x += sx; err += dy;
            EMIT();
        }
    } else {
        int err = dx - dy / 2;
        while (y != y1) {
            if (err >= 0) { x += sx; err -= dy; }
            y += sy; err += dx;
            EMIT();
        }
    }
#undef EMIT
    return n;
}

/* ---- SlidingWindowMax: util.hpp:199-253 (monotonic deque of indices) ------------------------------- */

typedef double (*in_fn)(const void* ctx, int i);

static void sliding_window_max(in_fn in, const void* ctx, int n, int w, double* out, int out_stride) {
    const int cap = (n > w ? n : w) + 1;
    int* q = (int*)malloc((size_t)cap * sizeof(int));
    int head = 0, tail = 0;                      /* deque = q[head .. tail) */
    int idx_in = 0, idx_out = 0;
    for (idx_in = 0; idx_in < w; ++idx_in) {     /* :218-226 first window (may read past the end) */
        while (tail > head && in(ctx, idx_in) >= in(ctx, q[tail - 1])) --tail;
        q[tail++] = idx_in;
    }
    for (; idx_in < n; ++idx_in) {               /* :232-247 */
        out[(size_t)(idx_out++) * out_stride] = in(ctx, q[head]);
        while (tail > head && q[head] <= idx_in - w) ++head;
        while (tail > head && in(ctx, idx_in) >= in(ctx, q[tail - 1])) --tail;
        q[tail++] = idx_in;
    }
    for (; idx_out < n; ++idx_out)               /* :250-252 repeat the last window */
        out[(size_t)idx_out * out_stride] = in(ctx, q[head]);
    free(q);
}

typedef struct { const double* p; int n; int stride; } line_ctx;
static double line_in(const void* c, int i) {
    const line_ctx* l = (const line_ctx*)c;
    return (i >= 0 && i < l->n) ? l->p[(size_t)i * l->stride] : 0.0;   /* Value(.., unknown) */
}

void orc_sliding_window_max(const double* in, int n, int w, double* out) {
    line_ctx c = {in, n, 1};
    sliding_window_max(line_in, &c, n, w, out, 1);
}

/* PrecomputeGridMap: src/my_lidar_graph_slam/mapping/grid_map_builder.cpp:518-536
 * = SlidingWindowMaxRow (:403-434, per column, sliding along y) into an intermediate map, then
 *   SlidingWindowMaxCol (:437-468, per row, sliding along x). */
void orc_precompute(const double* grid, int nx, int ny, int w, double* out) {
    double* tmp = (double*)calloc((size_t)nx * ny + 1, sizeof(double));
    for (int x = 0; x < nx; ++x) {
        line_ctx c = {grid + x, ny, nx};
        sliding_window_max(line_in, &c, ny, w, tmp + x, nx);
    }
    for (int y = 0; y < ny; ++y) {
        line_ctx c = {tmp + (size_t)y * nx, nx, 1};
        sliding_window_max(line_in, &c, nx, w, out + (size_t)y * nx, 1);
    }
    free(tmp);
}

/* PrecomputeGridMaps: grid_map_builder.cpp:471-495 (windows 1, 2, 4, ..., 2^height_max) */
void orc_pyramid(const double* grid, int nx, int ny, int height_max, double* out) {
    for (int h = 0, w = 1; h <= height_max; ++h, w <<= 1)
        orc_precompute(grid, nx, ny, w, out + (size_t)h * nx * ny);
}

/* ---- poses, scans -------------------------------------------------------------------------------- */

/* Compound: include/my_lidar_graph_slam/pose.hpp:150-161 */
void orc_compound(const double* a, const double* b, double* o) {
    const double s = sin(a[2]), c = cos(a[2]);
    const double x = c * b[0] - s * b[1] + a[0];
    const double y = s * b[0] + c * b[1] + a[1];
    o[0] = x; o[1] = y; o[2] = a[2] + b[2];
}

/* ScanData::HitPoint: include/my_lidar_graph_slam/sensor/sensor_data.hpp:162-173 */
static void hit_point(const double* sensor_pose, double range, double angle, double* hx, double* hy) {
    const double c = cos(sensor_pose[2] + angle);
    const double s = sin(sensor_pose[2] + angle);
    *hx = sensor_pose[0] + range * c;
    *hy = sensor_pose[1] + range * s;
}

/* GridMapBuilder::ComputeBoundingBoxAndScanPoints: grid_map_builder.cpp:335-380 */
int orc_hit_points(const double* robot_pose, const double* rel, int n, const double* angles,
                   const double* ranges, double scan_min_range, double scan_max_range,
                   double usable_min, double usable_max, double* sensor_pose, double* hit_xy,
                   double* bbox) {
    orc_compound(robot_pose, rel, sensor_pose);
    bbox[0] = bbox[2] = sensor_pose[0];
    bbox[1] = bbox[3] = sensor_pose[1];
    const double min_range = usable_min > scan_min_range ? usable_min : scan_min_range;   /* std::max */
    const double max_range = usable_max < scan_max_range ? usable_max : scan_max_range;   /* std::min */
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const double r = ranges[i];
        if (r >= max_range || r <= min_range) continue;
        double hx, hy;
        hit_point(sensor_pose, r, angles[i], &hx, &hy);
        hit_xy[2 * k] = hx; hit_xy[2 * k + 1] = hy; ++k;
        if (hx < bbox[0]) bbox[0] = hx;
        if (hy < bbox[1]) bbox[1] = hy;
        if (hx > bbox[2]) bbox[2] = hx;
        if (hy > bbox[3]) bbox[3] = hy;
    }
    return k;
}

/* ---- map geometry --------------------------------------------------------------------------------- */

/* GridMap::Resize: grid_map.hpp:652-711 (geometry only; contents are shifted by the caller) */
void orc_geometry_resize(const orc_geom* cur, double min_x, double min_y, double max_x, double max_y,
                         orc_geom* out, int* shift_x, int* shift_y) {
    const int p = cur->patch;
    const int px0 = cell_to_patch(world_to_cell(min_x, cur->min_x, cur->res), p);
    const int py0 = cell_to_patch(world_to_cell(min_y, cur->min_y, cur->res), p);
    const int px1 = cell_to_patch(world_to_cell(max_x, cur->min_x, cur->res), p);
    const int py1 = cell_to_patch(world_to_cell(max_y, cur->min_y, cur->res), p);
    const int npx = px1 - px0 + 1 > 0 ? px1 - px0 + 1 : 0;
    const int npy = py1 - py0 + 1 > 0 ? py1 - py0 + 1 : 0;
    *out = *cur;
    out->nx = npx * p;
    out->ny = npy * p;
    out->min_x = cur->min_x + (px0 * p) * cur->res;
    out->min_y = cur->min_y + (py0 * p) * cur->res;
    if (shift_x) *shift_x = px0 * p;
    if (shift_y) *shift_y = py0 * p;
}

/* GridMap::Expand: grid_map.hpp:715-736; returns 1 if the map was resized */
int orc_geometry_expand(const orc_geom* cur, double min_x, double min_y, double max_x, double max_y,
                        double enlarge_step, orc_geom* out, int* shift_x, int* shift_y) {
    const int ix0 = world_to_cell(min_x, cur->min_x, cur->res), iy0 = world_to_cell(min_y, cur->min_y, cur->res);
    const int ix1 = world_to_cell(max_x, cur->min_x, cur->res), iy1 = world_to_cell(max_y, cur->min_y, cur->res);
    const int in0 = ix0 >= 0 && ix0 < cur->nx && iy0 >= 0 && iy0 < cur->ny;
    const int in1 = ix1 >= 0 && ix1 < cur->nx && iy1 >= 0 && iy1 < cur->ny;
    if (in0 && in1) {
        *out = *cur;
        if (shift_x) *shift_x = 0;
        if (shift_y) *shift_y = 0;
        return 0;
    }
    double lo_x = cur->min_x + cur->res * 0, lo_y = cur->min_y + cur->res * 0;
    double hi_x = cur->min_x + cur->res * cur->nx, hi_y = cur->min_y + cur->res * cur->ny;
    lo_x = (min_x < lo_x) ? min_x - enlarge_step : lo_x;
    lo_y = (min_y < lo_y) ? min_y - enlarge_step : lo_y;
    hi_x = (max_x > hi_x) ? max_x + enlarge_step : hi_x;
    hi_y = (max_y > hi_y) ? max_y + enlarge_step : hi_y;
    orc_geometry_resize(cur, lo_x, lo_y, hi_x, hi_y, out, shift_x, shift_y);
    return 1;
}

/* ---- scan integration: grid_map_builder.cpp:159-186 (= :296-328) ----------------------------------- */

int orc_integrate_hits(double* grid, const orc_geom* g, const double* sensor_xy, int n,
                       const double* hit_xy, double p_hit, double p_miss) {
    const int sx = world_to_cell(sensor_xy[0], g->min_x, g->res);
    const int sy = world_to_cell(sensor_xy[1], g->min_y, g->res);
    int cap = 2 * (g->nx + g->ny) + 8, updates = 0;
    int* cells = (int*)malloc((size_t)cap * 2 * sizeof(int));
    for (int i = 0; i < n; ++i) {
        const int ex = world_to_cell(hit_xy[2 * i], g->min_x, g->res);
        const int ey = world_to_cell(hit_xy[2 * i + 1], g->min_y, g->res);
        /* ComputeMissedGridCellIndices (:384-396): Bresenham minus the last (hit) cell */
        const int m = orc_bresenham(sx, sy, ex, ey, cells, cap);
        if (m > cap) { free(cells); return -1; }
        for (int j = 0; j < m; ++j) {
            const int x = cells[2 * j], y = cells[2 * j + 1];
            if (x < 0 || x >= g->nx || y < 0 || y >= g->ny) { free(cells); return -1; }
            double* c = &grid[(size_t)y * g->nx + x];
            *c = orc_bayes_update(*c, j == m - 1 ? p_hit : p_miss);
        }
        updates += m;
    }
    free(cells);
    return updates;
}

/* ---- search step: scan_matcher_real_time_correlative.cpp:156-175 (= scan_matcher_branch_bound.cpp:178-197) */

static double search_step_theta(double res, int n, const double* ranges, double scan_range_max) {
    double max_r = ranges[0];
    for (int i = 1; i < n; ++i) if (ranges[i] > max_r) max_r = ranges[i];   /* std::max_element */
    const double max_range = max_r < scan_range_max ? max_r : scan_range_max;
    const double theta = res / max_range;
    return acos(1.0 - 0.5 * theta * theta);
}

/* ---- real-time correlative matcher: scan_matcher_real_time_correlative.cpp:50-256 ------------------- */

static double compute_score(const double* g, int nx, int ny, const int* idx, int n, int ox, int oy) {  /* :207-224 */
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += grid_value(g, nx, ny, idx[2 * i] + ox, idx[2 * i + 1] + oy);
    return sum;
}

int orc_rtcsm_match(const double* grid, const double* coarse, const orc_geom* g, int low_res,
                    double range_x, double range_y, double range_theta, double scan_range_max,
                    const double* init_pose, const double* rel, int n, const double* angles,
                    const double* ranges, double norm_threshold, orc_match* out) {
    double sensor[3];
    orc_compound(init_pose, rel, sensor);                                   /* :58-59 */
    const double step_x = g->res, step_y = g->res;
    const double step_t = search_step_theta(g->res, n, ranges, scan_range_max);
    const int win_x = (int)ceil(0.5 * range_x / step_x);                    /* :68-73 */
    const int win_y = (int)ceil(0.5 * range_y / step_y);
    const int win_t = (int)ceil(0.5 * range_theta / step_t);
    const double thr = norm_threshold * (double)(size_t)n;                 /* :76-77 */
    double score_max = thr;
    int bx = -win_x, by = -win_y, bt = -win_t;
    int* idx = (int*)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(int));
    long long scored = 0;
    for (int t = -win_t; t <= win_t; ++t) {                                 /* :88 */
        const double pose[3] = {sensor[0], sensor[1], sensor[2] + step_t * t};
        int k = 0;                                                          /* ComputeScanIndices :178-203 */
        for (int i = 0; i < n; ++i) {
            if (ranges[i] >= scan_range_max) continue;
            double hx, hy;
            hit_point(pose, ranges[i], angles[i], &hx, &hy);
            idx[2 * k] = world_to_cell(hx, g->min_x, g->res);
            idx[2 * k + 1] = world_to_cell(hy, g->min_y, g->res);
            ++k;
        }
        for (int x = -win_x; x <= win_x; x += low_res) {                    /* :98-99 */
            for (int y = -win_y; y <= win_y; y += low_res) {
                const double s = compute_score(coarse, g->nx, g->ny, idx, k, x, y);
                ++scored;
                if (s <= score_max) continue;                               /* :106-107 */
                for (int fx = x; fx < x + low_res; ++fx)                    /* EvaluateHighResolutionMap :227-256 */
                    for (int fy = y; fy < y + low_res; ++fy) {
                        const double fs = compute_score(grid, g->nx, g->ny, idx, k, fx, fy);
                        ++scored;
                        if (score_max < fs) { score_max = fs; bx = fx; by = fy; bt = t; }
                    }
            }
        }
    }
    free(idx);
    memset(out, 0, sizeof(*out));
    out->found = score_max > thr;                                           /* :118 */
    out->ix = bx; out->iy = by; out->it = bt;
    out->win_x = win_x; out->win_y = win_y; out->win_t = win_t;
    out->step_x = step_x; out->step_y = step_y; out->step_t = step_t;
    out->score = score_max;
    memcpy(out->sensor_pose, sensor, sizeof(sensor));
    out->best_sensor_pose[0] = sensor[0] + bx * step_x;                     /* :121-124 */
    out->best_sensor_pose[1] = sensor[1] + by * step_y;
    out->best_sensor_pose[2] = sensor[2] + bt * step_t;
    out->n_scored = scored;
    return 0;
}

/* ---- ScorePixelAccurate::Score: score_function_pixel_accurate.cpp:19-76 ---------------------------- */

double orc_pixel_accurate_score(const double* level, const orc_geom* g, double usable_min,
                                double usable_max, const double* sensor_pose, int n,
                                const double* angles, const double* ranges, double scan_min_range,
                                double scan_max_range) {
    double sum = 0.0;
    const double min_range = usable_min > scan_min_range ? usable_min : scan_min_range;
    const double max_range = usable_max < scan_max_range ? usable_max : scan_max_range;
    for (int i = 0; i < n; ++i) {
        const double r = ranges[i];
        if (r >= max_range || r <= min_range) continue;
        double hx, hy;
        hit_point(sensor_pose, r, angles[i], &hx, &hy);
        const double v = grid_value(level, g->nx, g->ny, world_to_cell(hx, g->min_x, g->res),
                                    world_to_cell(hy, g->min_y, g->res));
        if (v == 0.0) continue;                                             /* :52-53 */
        sum += v;
    }
    return sum;
}

/* ---- CostGreedyEndpoint (host tail of both matchers): cost_function_greedy_endpoint.cpp ------------- */

/* cost = {usableRangeMin, usableRangeMax, hitAndMissedDist, occupancyThreshold, kernelSize,
 *         scalingFactor, standardDeviation} in the constructor's order (:9-27). */
/* CostGreedyEndpoint::Cost: cost_function_greedy_endpoint.cpp:32-110; ScanData::HitAndMissedPoint:
 * sensor_data.hpp:177-198; GridMap::SquaredDistance: grid_map.hpp:895-902. */
double orc_cost_greedy_endpoint(const double* grid, const orc_geom* g, const double* cost,
                                const double* sensor_pose, int n, const double* angles,
                                const double* ranges, double scan_min_range, double scan_max_range) {
    const double usable_min = cost[0], usable_max = cost[1], dist = cost[2], occ_thr = cost[3];
    const int ks = (int)cost[4];
    const double scaling = cost[5], variance = cost[6] * cost[6];
    const double min_range = usable_min > scan_min_range ? usable_min : scan_min_range;   /* :39-42 */
    const double max_range = usable_max < scan_max_range ? usable_max : scan_max_range;
    double value = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = ranges[i];
        if (r >= max_range || r <= min_range) continue;                     /* :49-50 */
        const double c = cos(sensor_pose[2] + angles[i]);
        const double s = sin(sensor_pose[2] + angles[i]);
        const double hx = sensor_pose[0] + r * c, hy = sensor_pose[1] + r * s;
        const double mx = sensor_pose[0] + (r - dist) * c, my = sensor_pose[1] + (r - dist) * s;
        const int hix = world_to_cell(hx, g->min_x, g->res), hiy = world_to_cell(hy, g->min_y, g->res);
        const int mix = world_to_cell(mx, g->min_x, g->res), miy = world_to_cell(my, g->min_y, g->res);
        const double d0 = (ks + 1) * g->res;                                /* :64-66 */
        double min_sq = d0 * d0 + d0 * d0;
        for (int ky = -ks; ky <= ks; ++ky)
            for (int kx = -ks; kx <= ks; ++kx) {
                const double hv = grid_value(grid, g->nx, g->ny, hix + kx, hiy + ky);
                const double mv = grid_value(grid, g->nx, g->ny, mix + kx, miy + ky);
                if (hv == 0.0 || mv == 0.0) continue;                       /* :81-83 */
                if (hv < occ_thr || mv > occ_thr) continue;                 /* :89-91 */
                const double dx = kx * g->res, dy = ky * g->res;
                const double sq = dx * dx + dy * dy;
                if (sq < min_sq) min_sq = sq;                               /* std::min(sq, min_sq) */
            }
        value -= exp(-0.5 * min_sq / variance);                             /* :103 */
    }
    value *= scaling;                                                       /* :107 */
    return value;
}

/* The matchers' tail: scan_matcher_real_time_correlative.cpp:126-138 with ComputeGradient /
 * ComputeCovariance (cost_function_greedy_endpoint.cpp:113-171). cov is row-major 3x3. */
void orc_cost_tail(const double* grid, const orc_geom* g, const double* cost, const double* best_pose,
                   int n, const double* angles, const double* ranges, double scan_min_range,
                   double scan_max_range, double* normalized_cost, double* cov) {
    const double diff[3] = {g->res, g->res, 1e-2};                          /* :120-121 */
    double grad[3];
    *normalized_cost = orc_cost_greedy_endpoint(grid, g, cost, best_pose, n, angles, ranges,
                                                scan_min_range, scan_max_range) / (double)(size_t)n;
    for (int a = 0; a < 3; ++a) {
        double plus[3] = {best_pose[0], best_pose[1], best_pose[2]};
        double minus[3] = {best_pose[0], best_pose[1], best_pose[2]};
        for (int k = 0; k < 3; ++k) {                                       /* pose +/- delta, pose.hpp:60-74 */
            plus[k] = best_pose[k] + (k == a ? diff[a] : 0.0);
            minus[k] = best_pose[k] - (k == a ? diff[a] : 0.0);
        }
        const double d = orc_cost_greedy_endpoint(grid, g, cost, plus, n, angles, ranges, scan_min_range,
                                                  scan_max_range) -
                         orc_cost_greedy_endpoint(grid, g, cost, minus, n, angles, ranges, scan_min_range,
                                                  scan_max_range);
        grad[a] = 0.5 * d / diff[a];                                        /* :139-141 */
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[3 * i + j] = i == j ? grad[i] * grad[j] + 0.01 : grad[i] * grad[j];   /* :161-166 */
}

/* ---- grid-search matcher: scan_matcher_grid_search.cpp:45-114 --------------------------------------- */

/* Exhaustive search with ACCUMULATED offsets (dy += sy, ...), loop order y, x, theta, strict ">".
 * out->ix / iy / it are the loop counters of the winner, out->win_x / win_y / win_t the loop lengths. */
int orc_gs_match(const double* grid, const orc_geom* g, double range_x, double range_y, double range_theta,
                 double step_x, double step_y, double step_theta, double usable_min, double usable_max,
                 const double* init_pose, const double* rel, int n, const double* angles,
                 const double* ranges, double scan_min_range, double scan_max_range,
                 double norm_threshold, orc_match* out) {
    double sensor[3];
    orc_compound(init_pose, rel, sensor);                                   /* :55-56 */
    const double rx = range_x / 2.0, ry = range_y / 2.0, rt = range_theta / 2.0;   /* :59-61 */
    const double threshold = norm_threshold * (double)(size_t)n;            /* :67-68 */
    double score_max = threshold;
    memset(out, 0, sizeof(*out));
    out->ix = out->iy = out->it = -1;
    out->best_sensor_pose[0] = sensor[0]; out->best_sensor_pose[1] = sensor[1];
    out->best_sensor_pose[2] = sensor[2];                                   /* :71 */
    long long scored = 0;
    int ny = 0;
    for (double dy = -ry; dy <= ry; dy += step_y, ++ny) {                   /* :74-76 */
        int nx = 0;
        for (double dx = -rx; dx <= rx; dx += step_x, ++nx) {
            int nt = 0;
            for (double dt = -rt; dt <= rt; dt += step_theta, ++nt) {
                const double pose[3] = {sensor[0] + dx, sensor[1] + dy, sensor[2] + dt};
                const double s = orc_pixel_accurate_score(grid, g, usable_min, usable_max, pose, n, angles,
                                                          ranges, scan_min_range, scan_max_range);
                ++scored;
                if (s > score_max) {                                        /* :85-88 */
                    score_max = s;
                    out->ix = nx; out->iy = ny; out->it = nt;
                    out->best_sensor_pose[0] = pose[0]; out->best_sensor_pose[1] = pose[1];
                    out->best_sensor_pose[2] = pose[2];
                }
            }
            out->win_t = nt;
        }
        out->win_x = nx;
    }
    out->win_y = ny;
    out->found = score_max > threshold;                                     /* :94 */
    out->step_x = step_x; out->step_y = step_y; out->step_t = step_theta;
    out->score = orc_pixel_accurate_score(grid, g, usable_min, usable_max, out->best_sensor_pose, n, angles,
                                          ranges, scan_min_range, scan_max_range);
    out->sensor_pose[0] = sensor[0]; out->sensor_pose[1] = sensor[1]; out->sensor_pose[2] = sensor[2];
    out->n_scored = scored;
    return 0;
}

/* ---- branch-and-bound matcher: scan_matcher_branch_bound.cpp:47-163 -------------------------------- */

typedef struct { int x, y, t, h; } bb_node;

int orc_bb_match(const double* pyramid, const orc_geom* g, int height_max, double range_x,
                 double range_y, double range_theta, double scan_range_max, double usable_min,
                 double usable_max, const double* init_pose, const double* rel, int n,
                 const double* angles, const double* ranges, double scan_min_range,
                 double scan_max_range, double norm_threshold, orc_match* out) {
    double sensor[3];
    orc_compound(init_pose, rel, sensor);                                   /* :55-56 */
    const double step_x = g->res, step_y = g->res;
    const double step_t = search_step_theta(g->res, n, ranges, scan_range_max);
    const int win_x = (int)ceil(0.5 * range_x / step_x);                    /* :68-73 */
    const int win_y = (int)ceil(0.5 * range_y / step_y);
    const int win_t = (int)ceil(0.5 * range_theta / step_t);
    const double thr = norm_threshold * (double)(size_t)n;                 /* :75-76 */
    double score_max = thr;
    double best[3] = {sensor[0], sensor[1], sensor[2]};                     /* :80 */
    int bx = 0, by = 0, bt = 0, found_leaf = 0;
    const int win_size_max = 1 << height_max;
    size_t cap = 1024, sp = 0;
    bb_node* stack = (bb_node*)malloc(cap * sizeof(bb_node));
#define PUSH(X, Y, T, H) do { if (sp == cap) { cap *= 2; stack = (bb_node*)realloc(stack, cap * sizeof(bb_node)); } \
        stack[sp].x = (X); stack[sp].y = (Y); stack[sp].t = (T); stack[sp].h = (H); ++sp; } while (0)
    for (int x = -win_x; x <= win_x; x += win_size_max)                    /* :85-88 */
        for (int y = -win_y; y <= win_y; y += win_size_max)
            for (int t = -win_t; t <= win_t; ++t) PUSH(x, y, t, height_max);
    long long scored = 0;
    const size_t level_cells = (size_t)g->nx * g->ny;
    while (sp > 0) {                                                        /* :92 */
        const bb_node nd = stack[sp - 1];
        const double pose[3] = {sensor[0] + nd.x * step_x, sensor[1] + nd.y * step_y,
                                sensor[2] + nd.t * step_t};                 /* :96-99 */
        const double s = orc_pixel_accurate_score(pyramid + (size_t)nd.h * level_cells, g, usable_min,
                                                  usable_max, pose, n, angles, ranges, scan_min_range,
                                                  scan_max_range);
        ++scored;
        --sp;                                                               /* every branch pops */
        if (s <= score_max) continue;                                       /* :108-112 */
        if (nd.h == 0) {                                                    /* :115-121 */
            best[0] = pose[0]; best[1] = pose[1]; best[2] = pose[2];
            score_max = s; bx = nd.x; by = nd.y; bt = nd.t; found_leaf = 1;
        } else {                                                            /* :123-137 */
            const int h = nd.h - 1, w = 1 << h;
            PUSH(nd.x, nd.y, nd.t, h);
            PUSH(nd.x + w, nd.y, nd.t, h);
            PUSH(nd.x, nd.y + w, nd.t, h);
            PUSH(nd.x + w, nd.y + w, nd.t, h);
        }
    }
#undef PUSH
    free(stack);
    memset(out, 0, sizeof(*out));
    out->found = score_max > thr;                                           /* :143 */
    out->ix = bx; out->iy = by; out->it = bt;
    out->win_x = win_x; out->win_y = win_y; out->win_t = win_t;
    out->step_x = step_x; out->step_y = step_y; out->step_t = step_t;
    out->score = score_max;
    memcpy(out->sensor_pose, sensor, sizeof(sensor));
    memcpy(out->best_sensor_pose, best, sizeof(best));
    out->n_scored = scored;
    (void)found_leaf;
    return 0;
}
