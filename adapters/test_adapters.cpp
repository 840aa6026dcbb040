/* test_adapters.cpp -- runs the CUDA adapters side by side with the reference's own classes on a
 * synthetic office world and requires IDENTICAL outputs: every field of ScanMatchingSummary and
 * of each LoopDetectionResult (poses, cost, covariance), compared as bit patterns.
 * Built here against /root/reference (adapters/Makefile), run on the GPU box by
 * tests/test_gpu_adapters.py.  Exit code 0 = all identical. */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <type_traits>
#include <random>

#include "lgs_adapters/create_cuda_backends.hpp"
#include "lgs_adapters/grid_map_builder_cuda.hpp"
#include "lgs_adapters/loop_detector_branch_bound_cuda.hpp"
#include "lgs_adapters/loop_detector_real_time_correlative_cuda.hpp"
#include "lgs_adapters/scan_matcher_real_time_correlative_cuda.hpp"
#include "my_lidar_graph_slam/mapping/cost_function_greedy_endpoint.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_real_time_correlative.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_real_time_correlative.hpp"

using namespace MyLidarGraphSlam;
using namespace MyLidarGraphSlam::Mapping;

namespace {

/* 24 m x 24 m lattice of 4 m rooms with door gaps; analytic ray casting */
struct World {
    struct Seg { double x0, y0, x1, y1; };
    std::vector<Seg> segs;
    explicit World(unsigned seed) {
        std::mt19937 g(seed);
        std::uniform_real_distribution<double> u(0.4, 4.0 - 1.2 - 0.4);
        const double h = 12.0, room = 4.0, door = 1.2;
        segs = {{-h, -h, h, -h}, {h, -h, h, h}, {h, h, -h, h}, {-h, h, -h, -h}};
        for (int i = 1; i < 6; ++i) {
            const double c = -h + i * room;
            for (int j = 0; j < 6; ++j) {
                const double a0 = -h + j * room;
                double gp = a0 + u(g);
                segs.push_back({c, a0, c, gp}); segs.push_back({c, gp + door, c, a0 + room});
                gp = a0 + u(g);
                segs.push_back({a0, c, gp, c}); segs.push_back({gp + door, c, a0 + room, c});
            }
        }
    }
    double Cast(double x, double y, double a) const {
        const double c = std::cos(a), s = std::sin(a);
        double best = 1e9;
        for (const auto& sg : segs) {
            if (sg.y0 == sg.y1) {
                if (std::fabs(s) < 1e-12) continue;
                const double t = (sg.y0 - y) / s, xi = x + t * c;
                if (t > 1e-9 && xi >= std::min(sg.x0, sg.x1) && xi <= std::max(sg.x0, sg.x1)) best = std::min(best, t);
            } else {
                if (std::fabs(c) < 1e-12) continue;
                const double t = (sg.x0 - x) / c, yi = y + t * s;
                if (t > 1e-9 && yi >= std::min(sg.y0, sg.y1) && yi <= std::max(sg.y0, sg.y1)) best = std::min(best, t);
            }
        }
        return best;
    }
};

const RobotPose2D<double> kRelSensor(0.12, -0.03, 0.05);   /* non-trivial sensor mounting */

Sensor::ScanDataPtr<double> MakeScan(const World& w, const RobotPose2D<double>& robot, int n,
                                     std::mt19937& g, const double fov = 4.71238898038469) {
    std::normal_distribution<double> noise(0.0, 0.01);
    const RobotPose2D<double> sp = Compound(robot, kRelSensor);
    std::vector<double> ang(n), rng(n);
    for (int i = 0; i < n; ++i) {
        ang[i] = -fov / 2 + fov * i / (n - 1);
        rng[i] = std::min(30.0, std::max(0.02, w.Cast(sp.mX, sp.mY, sp.mTheta + ang[i]) + noise(g)));
    }
    return std::make_shared<Sensor::ScanData<double>>(
        "L", 0.0, robot, RobotPose2D<double>(0, 0, 0), kRelSensor, 0.02, 30.0, -fov / 2, fov / 2,
        std::move(ang), std::move(rng));
}

bool SameBits(double a, double b) { return std::memcmp(&a, &b, sizeof a) == 0; }
bool SamePose(const RobotPose2D<double>& a, const RobotPose2D<double>& b) {
    return SameBits(a.mX, b.mX) && SameBits(a.mY, b.mY) && SameBits(a.mTheta, b.mTheta);
}
bool SameMat(const Eigen::Matrix3d& a, const Eigen::Matrix3d& b) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) if (!SameBits(a(i, j), b(i, j))) return false;
    return true;
}

/* The smallest property-tree look-alike the factory templates need (Boost is absent here): every
 * key is looked up in one flat table, get_child returns the tree itself. */
struct FlatTree {
    std::map<std::string, std::string> kv;
    const FlatTree& get_child(const std::string&) const { return *this; }
    template <typename T> T conv(const std::string& v) const {
        if constexpr (std::is_same<T, std::string>::value) return v;
        else if constexpr (std::is_same<T, int>::value) return std::stoi(v);
        else return static_cast<T>(std::stod(v));
    }
    template <typename T> T get(const std::string& key) const { return conv<T>(kv.at(key)); }
    template <typename T> T get(const std::string& key, const T& def) const {
        const auto it = kv.find(key);
        return it == kv.end() ? def : conv<T>(it->second);
    }
    std::string get(const std::string& key, const char* def) const { return get<std::string>(key, def); }
    /* JSON arrays ("Devices": [0, 1]) as boost::property_tree shows them: a child whose elements have
     * empty keys; here the flat table holds them as one comma-separated string */
    struct Item { int v; template <typename T> T get_value() const { return static_cast<T>(v); } };
    struct ItemList {
        std::vector<std::pair<std::string, Item>> items;
        bool present = false;
        explicit operator bool() const { return present; }
        const std::vector<std::pair<std::string, Item>>& operator*() const { return items; }
    };
    ItemList get_child_optional(const std::string& key) const {
        ItemList list;
        const auto it = kv.find(key);
        if (it == kv.end()) return list;
        list.present = true;
        std::string tok;
        for (char ch : it->second + ",") {
            if (ch != ',') { tok += ch; continue; }
            if (!tok.empty()) list.items.push_back({std::string(), Item{std::stoi(tok)}});
            tok.clear();
        }
        return list;
    }
};

/* Geometry, patch allocation and every cell value, as bit patterns */
bool SameMap(const GridMapType& a, const GridMapType& b) {
    if (a.NumOfGridCellsX() != b.NumOfGridCellsX() || a.NumOfGridCellsY() != b.NumOfGridCellsY() ||
        a.NumOfPatchesX() != b.NumOfPatchesX() || a.NumOfPatchesY() != b.NumOfPatchesY() ||
        !SameBits(a.MinPos().mX, b.MinPos().mX) || !SameBits(a.MinPos().mY, b.MinPos().mY))
        return false;
    const int patch = a.PatchSize();
    for (int py = 0; py < a.NumOfPatchesY(); ++py)
        for (int px = 0; px < a.NumOfPatchesX(); ++px) {
            if (a.PatchIsAllocated(px, py) != b.PatchIsAllocated(px, py)) return false;
            if (!a.PatchIsAllocated(px, py)) continue;
            const auto* ca = a.PatchAt(px, py).Data();
            const auto* cb = b.PatchAt(px, py).Data();
            for (int k = 0; k < patch * patch; ++k)
                if (!SameBits(ca[k].Value(), cb[k].Value())) return false;
        }
    return true;
}

template <typename A, typename B>
bool SameBuilders(const A& a, const B& b) {
    if (a.LocalMaps().size() != b.LocalMaps().size() || !SameBits(a.AccumTravelDist(), b.AccumTravelDist()) ||
        a.LatestScanIdxMin() != b.LatestScanIdxMin() || a.LatestScanIdxMax() != b.LatestScanIdxMax() ||
        !SameMap(a.LatestMap(), b.LatestMap()))
        return false;
    for (std::size_t m = 0; m < a.LocalMaps().size(); ++m) {
        const LocalMapInfo& x = a.LocalMaps()[m];
        const LocalMapInfo& y = b.LocalMaps()[m];
        if (x.mIdx != y.mIdx || x.mPoseGraphNodeIdxMin != y.mPoseGraphNodeIdxMin ||
            x.mPoseGraphNodeIdxMax != y.mPoseGraphNodeIdxMax || x.mFinished != y.mFinished ||
            x.mPrecomputed != y.mPrecomputed || !SameMap(x.mMap, y.mMap))
            return false;
    }
    return true;
}

}  // namespace

/* ---- config C1: the front end's own frame loop (lidar_graph_slam_frontend.cpp:85-127) with the
 *      default launcher settings (launcher_settings_default.json:42-50, :175-185) on a 180-beam
 *      180-degree log with drifting odometry: reference classes vs the device pipeline (builder's
 *      device-resident latest map -> matcher -> device tail); poses feed back into the maps, so one
 *      differing bit anywhere compounds.  json: print one JSON line for bench.py (`extra.c1`). ---- */
int RunC1(const bool json) {
    std::mt19937 g(11);
    const World world(7);
    auto cost = std::make_shared<CostGreedyEndpoint>(0.01, 20.0, 0.075, 0.1, 1, 0.05, 1.0);
    const lgs_cost_params costParams { 0.01, 20.0, 0.075, 0.1, 1, 0.05, 1.0 };
    int failures = 0;
    const int numOfFrames = 150;
    std::vector<RobotPose2D<double>> truth, odom;
    for (int k = 0; k < numOfFrames; ++k) {
        truth.emplace_back(-9.9 + 0.09 * k * std::cos(0.004 * k), -9.8 + 0.5 * std::sin(0.05 * k), 0.1 * std::sin(0.03 * k));
        odom.emplace_back(truth[k].mX + 0.0015 * k, truth[k].mY - 0.001 * k, truth[k].mTheta + 0.0004 * k);
    }
    auto pgRef = std::make_shared<PoseGraph>(), pgCuda = std::make_shared<PoseGraph>();
    GridMapBuilder bRef(0.05, 64, 10, 20.0, 0.01, 20.0, 0.6, 0.45);
    GridMapBuilderCuda bCuda(0.05, 64, 10, 20.0, 0.01, 20.0, 0.6, 0.45, 0);
    bCuda.SetLazyHostMaps(true);     /* the matcher takes the device map; the host maps follow when they are read */
    ScanMatcherRealTimeCorrelative mRef(cost, 5, 0.2, 0.2, 0.5, 20.0);
    ScanMatcherRealTimeCorrelativeCuda mCuda(cost, 5, 0.2, 0.2, 0.5, 20.0, 0);
    mCuda.UseDeviceCost(costParams);
    double msRef = 0.0, msCuda = 0.0, msMatch = 0.0, warm[4] = { 0.0, 0.0, 0.0, 0.0 };
    int bad = 0, lost = 0;
    for (int k = 0; k < numOfFrames; ++k) {
        if (k == 5) for (int j = 0; j < 4; ++j) warm[j] = bCuda.TimingsMs()[j];   /* allocations settle first */
        const auto scan = MakeScan(world, truth[k], 180, g, 3.14159265358979323846);
        const auto t0 = std::chrono::steady_clock::now();
        if (k == 0) {
            pgRef->AppendNode(truth[0], scan);
        } else {
            const RobotPose2D<double> initial =
                Compound(pgRef->LatestNode().Pose(), InverseCompound(odom[k - 1], odom[k]));
            const ScanMatchingQuery query(GridMapType(bRef.LatestMap()), scan, initial);
            const ScanMatchingSummary a = mRef.OptimizePose(query);
            lost += !a.mPoseFound;
            pgRef->AppendNode(a.mEstimatedPose, scan);
        }
        bRef.AppendScan(pgRef);
        const auto t1 = std::chrono::steady_clock::now();
        if (k == 0) {
            pgCuda->AppendNode(truth[0], scan);
        } else {
            const RobotPose2D<double> initial =
                Compound(pgCuda->LatestNode().Pose(), InverseCompound(odom[k - 1], odom[k]));
            const ScanMatchingSummary b = mCuda.OptimizePose(bCuda.DeviceLatestMap(), scan, initial,
                                                             std::numeric_limits<double>::min());
            pgCuda->AppendNode(b.mEstimatedPose, scan);
        }
        if (k >= 5) msMatch += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
        bCuda.AppendScan(pgCuda);
        const auto t2 = std::chrono::steady_clock::now();
        if (k >= 5) { msRef += std::chrono::duration<double, std::milli>(t1 - t0).count();
                      msCuda += std::chrono::duration<double, std::milli>(t2 - t1).count(); }
        bad += !SamePose(pgRef->LatestNode().Pose(), pgCuda->LatestNode().Pose());
    }
    const bool maps = SameBuilders(bRef, bCuda);
    const RobotPose2D<double>& last = pgCuda->LatestNode().Pose();
    const double err = std::hypot(last.mX - truth.back().mX, last.mY - truth.back().mY);
    std::printf("C1 front-end loop, %d frames of 180 beams (default settings): poses %s, maps %s; per frame reference "
                "%.2f ms, cuda %.2f ms; end position error %.3f m (odometry alone %.3f m), %d lost\n", numOfFrames,
                bad ? "MISMATCH" : "IDENTICAL", maps ? "IDENTICAL" : "MISMATCH", msRef / (numOfFrames - 5),
                msCuda / (numOfFrames - 5), err, std::hypot(0.0015, 0.001) * (numOfFrames - 1), lost);
    const double* tm = bCuda.TimingsMs();
    const int steady = numOfFrames - 5;
    std::printf("C1 cuda split per frame after 5 warm-up frames: match %.2f ms; builder: integrate %.2f ms, download %.2f ms, "
                "host write-back %.2f ms (local map %dx%d, latest map %dx%d)\n", msMatch / steady,
                (tm[1] - warm[1]) / steady, (tm[2] - warm[2]) / steady, (tm[3] - warm[3]) / steady,
                bCuda.LocalMaps().back().mMap.NumOfGridCellsX(),
                bCuda.LocalMaps().back().mMap.NumOfGridCellsY(), bCuda.LatestMap().NumOfGridCellsX(),
                bCuda.LatestMap().NumOfGridCellsY());
    failures += (bad != 0) + !maps + (lost != 0);
    if (json)
        std::printf("{\"frames\": %d, \"beams\": 180, \"frames_per_s\": %.1f, \"ref_frames_per_s\": %.1f, "
                    "\"ms_per_frame\": %.3f, \"ref_ms_per_frame\": %.3f, \"ms_match\": %.3f, \"ms_integrate\": %.3f, "
                    "\"ms_download\": %.3f, \"ms_host_writeback\": %.3f, \"identical\": %s, \"lost\": %d}\n",
                    numOfFrames, 1e3 * steady / msCuda, 1e3 * steady / msRef, msCuda / steady, msRef / steady,
                    msMatch / steady, (tm[1] - warm[1]) / steady, (tm[2] - warm[2]) / steady, (tm[3] - warm[3]) / steady,
                    (bad == 0 && maps) ? "true" : "false", lost);
    return failures;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--c1-json")
        return RunC1(true) ? 1 : 0;
    std::mt19937 g(5);
    const World world(7);
    auto poseGraph = std::make_shared<PoseGraph>();
    GridMapBuilder builder(0.05, 64, 10, 6.0, 0.01, 20.0, 0.6, 0.45);   /* new local map every 6 m */
    GridMapBuilderCuda builderCuda(0.05, 64, 10, 6.0, 0.01, 20.0, 0.6, 0.45, 0);
    int failures = 0;
    /* drive around inside one room row, through doors where they line up */
    std::vector<RobotPose2D<double>> path;
    for (int k = 0; k < 60; ++k)
        path.emplace_back(-9.8 + 0.33 * k * std::cos(0.02 * k), -9.9 + 0.9 * std::sin(0.21 * k), 0.03 * k);
    {   /* ---- grid map builder: every frame, then a loop closure, then the global map ---- */
        int frame = 0, bad = 0;
        double msRef = 0.0, msCuda = 0.0, worstCuda = 0.0;
        int worstFrame = -1;
        for (const auto& p : path) {
            poseGraph->AppendNode(p, MakeScan(world, p, 541, g));
            const auto t0 = std::chrono::steady_clock::now();
            const bool c1 = builder.AppendScan(poseGraph);
            const auto t1 = std::chrono::steady_clock::now();
            const bool c2 = builderCuda.AppendScan(poseGraph);
            const auto t2 = std::chrono::steady_clock::now();
            if (frame >= 5) {
                const double dc = std::chrono::duration<double, std::milli>(t2 - t1).count();
                msRef += std::chrono::duration<double, std::milli>(t1 - t0).count();
                msCuda += dc;
                if (dc > worstCuda) { worstCuda = dc; worstFrame = frame; }
            }
            if (c1 != c2 || !SameBuilders(builder, builderCuda)) {
                if (bad++ < 5) std::printf("builder frame %d: MISMATCH\n", frame);
            }
            ++frame;
        }
        std::printf("builder: %d frames, %zu local maps, %lld cell updates on the device %s\n", frame,
                    builderCuda.LocalMaps().size(), builderCuda.NumOfCellUpdates(), bad ? "MISMATCH" : "IDENTICAL");
        failures += bad != 0;
        std::printf("builder AppendScan per frame (541 beams, after 5 warm-up frames): reference %.2f ms, cuda %.2f ms "
                    "(worst cuda frame %d: %.2f ms, mean without it %.2f ms)\n",
                    msRef / (frame - 5), msCuda / (frame - 5), worstFrame, worstCuda,
                    (msCuda - worstCuda) / (frame - 6));
        /* pretend a loop closure moved the nodes (smooth drift correction), rebuild everything */
        auto corrected = std::make_shared<PoseGraph>(*poseGraph);
        for (int i = 0; i < static_cast<int>(corrected->Nodes().size()); ++i) {
            RobotPose2D<double>& q = corrected->NodeAt(i).Pose();
            q.mX += 0.004 * i; q.mY -= 0.003 * i; q.mTheta += 0.0005 * i;
        }
        GridMapBuilder refCopy(builder);
        refCopy.LocalMapAt(0).mPrecomputed = true;
        builderCuda.LocalMapAt(0).mPrecomputed = true;
        (void)builderCuda.ConstructGlobalMap(poseGraph);   /* grows the device workspace once (not timed) */
        const auto a0 = std::chrono::steady_clock::now();
        refCopy.AfterLoopClosure(corrected);
        const auto a1 = std::chrono::steady_clock::now();
        builderCuda.AfterLoopClosure(corrected);
        const auto a2 = std::chrono::steady_clock::now();
        bool ok = SameBuilders(refCopy, builderCuda);
        std::printf("builder AfterLoopClosure (%d nodes): reference %.1f ms, cuda %.1f ms %s\n", frame,
                    std::chrono::duration<double, std::milli>(a1 - a0).count(),
                    std::chrono::duration<double, std::milli>(a2 - a1).count(), ok ? "IDENTICAL" : "MISMATCH");
        failures += !ok;
        /* keep mapping after the closure: the device mirror of the current local map must follow */
        for (int k = 0; k < 4; ++k) {
            const RobotPose2D<double> p(path.back().mX + 0.3 * (k + 1), path.back().mY + 0.1 * k, path.back().mTheta);
            corrected->AppendNode(p, MakeScan(world, p, 541, g));
            refCopy.AppendScan(corrected);
            builderCuda.AppendScan(corrected);
        }
        ok = SameBuilders(refCopy, builderCuda);
        std::printf("builder frames after the closure: %s\n", ok ? "IDENTICAL" : "MISMATCH");
        failures += !ok;
        ok = SameMap(refCopy.ConstructGlobalMap(corrected), builderCuda.ConstructGlobalMap(corrected));
        std::printf("builder ConstructGlobalMap: %s\n", ok ? "IDENTICAL" : "MISMATCH");
        failures += !ok;
    }
    std::printf("local maps: %zu, latest map %dx%d\n", builder.LocalMaps().size(),
                builder.LatestMap().NumOfGridCellsX(), builder.LatestMap().NumOfGridCellsY());
    /* cost function exactly as slam_launcher.cpp:60-72 builds it from the default settings */
    auto cost = std::make_shared<CostGreedyEndpoint>(0.01, 20.0, 0.075, 0.1, 1, 0.05, 1.0);

    /* the same arguments for the device evaluation of the tail (lgs_cost_tail) */
    const lgs_cost_params costParams { 0.01, 20.0, 0.075, 0.1, 1, 0.05, 1.0 };

    /* ---- front-end matcher: tail on the host (reference code), then on the device ---- */
    for (int deviceCost = 0; deviceCost < 2; ++deviceCost) {
        ScanMatcherRealTimeCorrelative ref(cost, 5, 1.0, 1.0, 0.6, 20.0);
        ScanMatcherRealTimeCorrelativeCuda gpu(cost, 5, 1.0, 1.0, 0.6, 20.0, 0);
        if (deviceCost) gpu.UseDeviceCost(costParams);
        std::uniform_real_distribution<double> dxy(-0.3, 0.3), dth(-0.15, 0.15);
        double msMatchRef = 0.0, msMatchCuda = 0.0;
        for (int k = 0; k < 6; ++k) {
            const RobotPose2D<double> truth = path[50 + k];
            const auto scan = MakeScan(world, truth, 541, g);
            const RobotPose2D<double> init(truth.mX + dxy(g), truth.mY + dxy(g), truth.mTheta + dth(g));
            ScanMatchingQuery q1(GridMapType(builder.LatestMap()), scan, init);
            ScanMatchingQuery q2(GridMapType(builder.LatestMap()), scan, init);
            const auto m0 = std::chrono::steady_clock::now();
            const ScanMatchingSummary a = ref.OptimizePose(q1);
            const auto m1 = std::chrono::steady_clock::now();
            const ScanMatchingSummary b = gpu.OptimizePose(q2);
            const auto m2 = std::chrono::steady_clock::now();
            if (k > 0) { msMatchRef += std::chrono::duration<double, std::milli>(m1 - m0).count();
                         msMatchCuda += std::chrono::duration<double, std::milli>(m2 - m1).count(); }
            const bool ok = a.mPoseFound == b.mPoseFound && SameBits(a.mNormalizedCost, b.mNormalizedCost) &&
                            SamePose(a.mInitialPose, b.mInitialPose) && SamePose(a.mEstimatedPose, b.mEstimatedPose) &&
                            SameMat(a.mEstimatedCovariance, b.mEstimatedCovariance);
            std::printf("rtcsm %d: found %d/%d pose (%.4f %.4f %.4f) idx (%d %d %d) score %.6f %s\n", k,
                        a.mPoseFound, b.mPoseFound, b.mEstimatedPose.mX, b.mEstimatedPose.mY, b.mEstimatedPose.mTheta,
                        gpu.LastResult().ix, gpu.LastResult().iy, gpu.LastResult().it, gpu.LastResult().score,
                        ok ? "IDENTICAL" : "MISMATCH");
            failures += !ok;
        }
        std::printf("rtcsm OptimizePose(query) per call (map upload + coarse map + sweep + %s tail): "
                    "reference %.2f ms, cuda %.2f ms\n", deviceCost ? "device" : "host", msMatchRef / 5, msMatchCuda / 5);
    }

    /* ---- front-end matcher fed from the builder's device-resident latest map (no host round trip) ---- */
    {
        auto pg = std::make_shared<PoseGraph>();
        GridMapBuilderCuda live(0.05, 64, 10, 6.0, 0.01, 20.0, 0.6, 0.45, 0);
        ScanMatcherRealTimeCorrelative ref(cost, 5, 1.0, 1.0, 0.6, 20.0);
        ScanMatcherRealTimeCorrelativeCuda gpu(cost, 5, 1.0, 1.0, 0.6, 20.0, 0);
        gpu.UseDeviceCost(costParams);
        std::uniform_real_distribution<double> dxy(-0.3, 0.3), dth(-0.15, 0.15);
        double msRef = 0.0, msCuda = 0.0;
        int bad = 0, frames = 0;
        for (int k = 0; k < 16; ++k) {
            pg->AppendNode(path[k], MakeScan(world, path[k], 541, g));
            live.AppendScan(pg);
            if (live.DeviceLatestMap() == nullptr) { ++bad; continue; }
            const RobotPose2D<double> truth = path[k + 1];
            const auto scan = MakeScan(world, truth, 541, g);
            const RobotPose2D<double> init(truth.mX + dxy(g), truth.mY + dxy(g), truth.mTheta + dth(g));
            ScanMatchingQuery q1(GridMapType(live.LatestMap()), scan, init);
            const auto m0 = std::chrono::steady_clock::now();
            const ScanMatchingSummary a = ref.OptimizePose(q1);
            const auto m1 = std::chrono::steady_clock::now();
            const ScanMatchingSummary b = gpu.OptimizePose(live.DeviceLatestMap(), scan, init,
                                                           std::numeric_limits<double>::min());
            const auto m2 = std::chrono::steady_clock::now();
            if (k >= 4) { msRef += std::chrono::duration<double, std::milli>(m1 - m0).count();
                          msCuda += std::chrono::duration<double, std::milli>(m2 - m1).count(); ++frames; }
            bad += !(a.mPoseFound == b.mPoseFound && SameBits(a.mNormalizedCost, b.mNormalizedCost) &&
                     SamePose(a.mInitialPose, b.mInitialPose) && SamePose(a.mEstimatedPose, b.mEstimatedPose) &&
                     SameMat(a.mEstimatedCovariance, b.mEstimatedCovariance));
        }
        std::printf("rtcsm against the builder's device-resident latest map (16 frames): %s; per call reference %.2f ms, "
                    "cuda %.3f ms\n", bad ? "MISMATCH" : "IDENTICAL", msRef / frames, msCuda / frames);
        failures += bad != 0;
        (void)live.ConstructGlobalMap(pg);
        if (live.DeviceLatestMap() != nullptr) { std::printf("device latest map must be invalid after ConstructGlobalMap\n"); ++failures; }
    }

    failures += RunC1(false);

    /* ---- loop detector ---- */
    {
        auto score = std::make_shared<ScorePixelAccurate>(0.01, 20.0);
        auto bb = std::make_shared<ScanMatcherBranchBound>(score, cost, 6, 2.0, 2.0, 1.0, 20.0);
        LoopDetectorBranchBound ref(bb, 0.6);
        LoopDetectorBranchBoundCuda gpu(0.01, 20.0, cost, 6, 2.0, 2.0, 1.0, 20.0, 0.6, 0);
        std::uniform_real_distribution<double> dxy(-0.5, 0.5), dth(-0.25, 0.25);
        for (int round = 0; round < 2; ++round) {   /* second round hits the device pyramid cache */
            if (round == 1) gpu.UseDeviceCost(costParams);   /* ... and evaluates the covariances on the device */
            LoopDetectionQueryVector q1, q2;
            for (size_t m = 0; m + 1 < builder.LocalMaps().size() && m < 3; ++m) {
                LocalMapInfo info = builder.LocalMapAt(static_cast<int>(m));
                info.mFinished = true;
                std::vector<PoseGraph::Node> n1, n2;
                for (int j = 0; j < 3; ++j) {   /* revisit poses inside this local map, perturbed */
                    const int idx = info.mPoseGraphNodeIdxMin + 2 + 3 * j + round;
                    const RobotPose2D<double> truth = poseGraph->NodeAt(idx).Pose();
                    const auto scan = MakeScan(world, truth, 541, g);
                    const RobotPose2D<double> pert(truth.mX + dxy(g), truth.mY + dxy(g), truth.mTheta + dth(g));
                    n1.emplace_back(1000 + j, pert, scan);
                    n2.emplace_back(1000 + j, pert, scan);
                }
                const PoseGraph::Node& mapNode = poseGraph->NodeAt(info.mPoseGraphNodeIdxMin);
                q1.emplace_back(std::move(n1), info, mapNode);
                q2.emplace_back(std::move(n2), info, mapNode);
            }
            LoopDetectionResultVector r1, r2;
            const auto d0 = std::chrono::steady_clock::now();
            ref.Detect(q1, r1);
            const auto d1 = std::chrono::steady_clock::now();
            gpu.Detect(q2, r2);
            const auto d2 = std::chrono::steady_clock::now();
            long long fix = 0, replay = 0, scored = 0;          /* n_scored is per pair (bb_count_nodes) */
            for (const auto& r : gpu.LastResults()) { fix += r.n_fixups; replay += r.exact_replay; scored += r.n_scored; }
            std::printf("loop round %d Detect: reference %.1f ms, cuda %.1f ms (%lld nodes scored, %lld host fix-ups, "
                        "%lld CPU-order replays)\n", round,
                        std::chrono::duration<double, std::milli>(d1 - d0).count(),
                        std::chrono::duration<double, std::milli>(d2 - d1).count(),
                        scored, fix, replay);
            bool ok = r1.size() == r2.size();
            for (size_t i = 0; ok && i < r1.size(); ++i)
                ok = SamePose(r1[i].mRelativePose, r2[i].mRelativePose) &&
                     SamePose(r1[i].mStartNodePose, r2[i].mStartNodePose) &&
                     r1[i].mStartNodeIdx == r2[i].mStartNodeIdx && r1[i].mEndNodeIdx == r2[i].mEndNodeIdx &&
                     SameMat(r1[i].mEstimatedCovMat, r2[i].mEstimatedCovMat);
            for (auto& q : q2) ok = ok && q.mLocalMapInfo.mPrecomputed;
            std::printf("loop round %d: %zu queries, %zu pairs, loops ref %zu / cuda %zu %s\n", round, q1.size(),
                        gpu.LastResults().size(), r1.size(), r2.size(), ok ? "IDENTICAL" : "MISMATCH");
            failures += !ok;
            if (r1.empty()) { std::printf("expected at least one detected loop\n"); ++failures; }
        }
    }
    /* ---- loop detector over every GPU of the box: local map i on device i mod G, one
     * process, one host thread per device, records exchanged on the devices.  Also a Detect() call in
     * which two queries name the SAME local map (each query owns a copy of LocalMapInfo whose
     * mPrecomputed flag is still false: the pyramid built for the first must survive the second). ---- */
    {
        const int numOfDevices = lgs_device_count();
        std::vector<int> devices;
        for (int k = 0; k < numOfDevices && k < 8; ++k) devices.push_back(k);
        auto score = std::make_shared<ScorePixelAccurate>(0.01, 20.0);
        auto bb = std::make_shared<ScanMatcherBranchBound>(score, cost, 6, 2.0, 2.0, 1.0, 20.0);
        LoopDetectorBranchBound ref(bb, 0.6);
        LoopDetectorBranchBoundCuda gpu(0.01, 20.0, cost, 6, 2.0, 2.0, 1.0, 20.0, 0.6, devices);
        gpu.UseDeviceCost(costParams);
        std::uniform_real_distribution<double> dxy(-0.5, 0.5), dth(-0.25, 0.25);
        LoopDetectionQueryVector q1, q2;
        const size_t numOfMaps = std::min<size_t>(builder.LocalMaps().size() - 1, 4);
        for (size_t k = 0; k < numOfMaps + 1; ++k) {
            const size_t m = k < numOfMaps ? k : 0;          /* the last query names local map 0 again */
            LocalMapInfo info = builder.LocalMapAt(static_cast<int>(m));
            info.mFinished = true;
            std::vector<PoseGraph::Node> n1, n2;
            for (int j = 0; j < 2; ++j) {
                const int idx = info.mPoseGraphNodeIdxMin + 1 + 4 * j + static_cast<int>(k);
                const RobotPose2D<double> truth = poseGraph->NodeAt(std::min(idx, info.mPoseGraphNodeIdxMax)).Pose();
                const auto scan = MakeScan(world, truth, 541, g);
                const RobotPose2D<double> pert(truth.mX + dxy(g), truth.mY + dxy(g), truth.mTheta + dth(g));
                n1.emplace_back(3000 + 10 * static_cast<int>(k) + j, pert, scan);
                n2.emplace_back(3000 + 10 * static_cast<int>(k) + j, pert, scan);
            }
            const PoseGraph::Node& mapNode = poseGraph->NodeAt(info.mPoseGraphNodeIdxMin);
            q1.emplace_back(std::move(n1), info, mapNode);
            q2.emplace_back(std::move(n2), info, mapNode);
        }
        LoopDetectionResultVector r1, r2;
        ref.Detect(q1, r1);
        const auto d1 = std::chrono::steady_clock::now();
        gpu.Detect(q2, r2);
        const auto d2 = std::chrono::steady_clock::now();
        LoopDetectionResultVector r3;
        gpu.Detect(q2, r3);                                   /* cached pyramids on every device */
        const auto d3 = std::chrono::steady_clock::now();
        bool ok = r1.size() == r2.size() && r1.size() == r3.size();
        for (size_t i = 0; ok && i < r1.size(); ++i)
            ok = SamePose(r1[i].mRelativePose, r2[i].mRelativePose) && SamePose(r1[i].mRelativePose, r3[i].mRelativePose) &&
                 r1[i].mStartNodeIdx == r2[i].mStartNodeIdx && r1[i].mEndNodeIdx == r2[i].mEndNodeIdx &&
                 SameMat(r1[i].mEstimatedCovMat, r2[i].mEstimatedCovMat) && SameMat(r1[i].mEstimatedCovMat, r3[i].mEstimatedCovMat);
        for (auto& q : q2) ok = ok && q.mLocalMapInfo.mPrecomputed;
        std::printf("loop detector on %zu device(s): %zu queries (one local map twice), loops ref %zu / cuda %zu, "
                    "Detect %.1f ms then %.1f ms: %s\n", devices.size(), q1.size(), r1.size(), r2.size(),
                    std::chrono::duration<double, std::milli>(d2 - d1).count(),
                    std::chrono::duration<double, std::milli>(d3 - d2).count(), ok ? "IDENTICAL" : "MISMATCH");
        failures += !ok;
        if (r1.empty()) { std::printf("expected at least one detected loop\n"); ++failures; }
    }
    /* ---- loop detector built on the correlative matcher ---- */
    {
        auto refMatcher = std::make_shared<ScanMatcherRealTimeCorrelative>(cost, 5, 1.5, 1.5, 0.6, 20.0);
        auto gpuMatcher = std::make_shared<ScanMatcherRealTimeCorrelativeCuda>(cost, 5, 1.5, 1.5, 0.6, 20.0, 0);
        gpuMatcher->UseDeviceCost(costParams);
        LoopDetectorRealTimeCorrelative ref(refMatcher, 0.55);
        LoopDetectorRealTimeCorrelativeCuda gpu(gpuMatcher, 0.55);
        std::uniform_real_distribution<double> dxy(-0.5, 0.5), dth(-0.2, 0.2);
        LoopDetectionQueryVector q1, q2;
        for (size_t m = 0; m + 1 < builder.LocalMaps().size() && m < 3; ++m) {
            LocalMapInfo info = builder.LocalMapAt(static_cast<int>(m));
            info.mFinished = true;
            std::vector<PoseGraph::Node> n1, n2;
            for (int j = 0; j < 4; ++j) {
                const int idx = info.mPoseGraphNodeIdxMin + 1 + 3 * j;
                const RobotPose2D<double> truth = poseGraph->NodeAt(idx).Pose();
                const auto scan = MakeScan(world, truth, 541, g);
                /* the last node is far off: its score stays below the threshold (no edge emitted) */
                const double far = j == 3 ? 3.0 : 0.0;
                const RobotPose2D<double> pert(truth.mX + dxy(g) + far, truth.mY + dxy(g) - far, truth.mTheta + dth(g));
                n1.emplace_back(2000 + j, pert, scan);
                n2.emplace_back(2000 + j, pert, scan);
            }
            const PoseGraph::Node& mapNode = poseGraph->NodeAt(info.mPoseGraphNodeIdxMin);
            q1.emplace_back(std::move(n1), info, mapNode);
            q2.emplace_back(std::move(n2), info, mapNode);
        }
        LoopDetectionResultVector r1, r2;
        const auto c0 = std::chrono::steady_clock::now();
        ref.Detect(q1, r1);
        const auto c1 = std::chrono::steady_clock::now();
        gpu.Detect(q2, r2);
        const auto c2 = std::chrono::steady_clock::now();
        std::printf("correlative loop detector Detect (device tail): reference %.1f ms, cuda %.1f ms\n",
                    std::chrono::duration<double, std::milli>(c1 - c0).count(),
                    std::chrono::duration<double, std::milli>(c2 - c1).count());
        bool ok = r1.size() == r2.size();
        for (size_t i = 0; ok && i < r1.size(); ++i)
            ok = SamePose(r1[i].mRelativePose, r2[i].mRelativePose) &&
                 SamePose(r1[i].mStartNodePose, r2[i].mStartNodePose) &&
                 r1[i].mStartNodeIdx == r2[i].mStartNodeIdx && r1[i].mEndNodeIdx == r2[i].mEndNodeIdx &&
                 SameMat(r1[i].mEstimatedCovMat, r2[i].mEstimatedCovMat);
        for (auto& q : q2) ok = ok && q.mLocalMapInfo.mPrecomputed;
        size_t nodes = 0;
        for (auto& q : q1) nodes += q.mPoseGraphNodes.size();
        std::printf("correlative loop detector: %zu queries, %zu nodes, loops ref %zu / cuda %zu %s\n", q1.size(), nodes,
                    r1.size(), r2.size(), ok ? "IDENTICAL" : "MISMATCH");
        failures += !ok;
        if (r1.empty() || r1.size() == nodes) { std::printf("expected some but not all nodes to close a loop\n"); ++failures; }
    }
    /* ---- grid-search matcher and the loop detector built on it (small windows: the CPU search projects
     *      the whole scan once per hypothesis) ---- */
    {
        auto score = std::make_shared<ScorePixelAccurate>(0.01, 20.0);
        auto refMatcher = std::make_shared<ScanMatcherGridSearch>(score, cost, 0.5, 0.4, 0.1, 0.05, 0.05, 0.01);
        auto gpuMatcher = std::make_shared<ScanMatcherGridSearchCuda>(0.01, 20.0, cost, 0.5, 0.4, 0.1, 0.05, 0.05, 0.01, 0);
        LoopDetectorGridSearch ref(refMatcher, 0.5);
        LoopDetectorGridSearchCuda gpu(gpuMatcher, 0.5);
        std::uniform_real_distribution<double> dxy(-0.15, 0.15), dth(-0.03, 0.03);
        for (int round = 0; round < 2; ++round) {      /* host tail, then device tail */
            if (round == 1) gpuMatcher->UseDeviceCost(costParams);
            /* the matcher itself, including a match that stays below the threshold */
            int bad = 0;
            for (int k = 0; k < 3; ++k) {
                const RobotPose2D<double> truth = path[52 + k];
                const auto scan = MakeScan(world, truth, 541, g);
                const double far = k == 2 ? 4.0 : 0.0;
                const RobotPose2D<double> init(truth.mX + dxy(g) + far, truth.mY + dxy(g), truth.mTheta + dth(g));
                const ScanMatchingSummary a = refMatcher->OptimizePose(builder.LatestMap(), scan, init, 0.55);
                const ScanMatchingSummary b = gpuMatcher->OptimizePose(builder.LatestMap(), scan, init, 0.55);
                bad += !(a.mPoseFound == b.mPoseFound && SameBits(a.mNormalizedCost, b.mNormalizedCost) &&
                         SamePose(a.mInitialPose, b.mInitialPose) && SamePose(a.mEstimatedPose, b.mEstimatedPose) &&
                         SameMat(a.mEstimatedCovariance, b.mEstimatedCovariance));
                if (k == 2 && b.mPoseFound) { std::printf("grid search: the far-off match must not be found\n"); ++bad; }
            }
            LoopDetectionQueryVector q1, q2;
            for (size_t m = 0; m + 1 < builder.LocalMaps().size() && m < 2; ++m) {
                LocalMapInfo info = builder.LocalMapAt(static_cast<int>(m));
                info.mFinished = true;
                std::vector<PoseGraph::Node> n1, n2;
                for (int j = 0; j < 3; ++j) {
                    const int idx = info.mPoseGraphNodeIdxMin + 2 + 3 * j + round;
                    const RobotPose2D<double> truth = poseGraph->NodeAt(idx).Pose();
                    const auto scan = MakeScan(world, truth, 541, g);
                    const double far = j == 2 ? 3.0 : 0.0;
                    const RobotPose2D<double> pert(truth.mX + dxy(g) + far, truth.mY + dxy(g) - far, truth.mTheta + dth(g));
                    n1.emplace_back(3000 + j, pert, scan);
                    n2.emplace_back(3000 + j, pert, scan);
                }
                const PoseGraph::Node& mapNode = poseGraph->NodeAt(info.mPoseGraphNodeIdxMin);
                q1.emplace_back(std::move(n1), info, mapNode);
                q2.emplace_back(std::move(n2), info, mapNode);
            }
            LoopDetectionResultVector r1, r2;
            const auto s0 = std::chrono::steady_clock::now();
            ref.Detect(q1, r1);
            const auto s1 = std::chrono::steady_clock::now();
            gpu.Detect(q2, r2);
            const auto s2 = std::chrono::steady_clock::now();
            bool ok = bad == 0 && r1.size() == r2.size();
            for (size_t i = 0; ok && i < r1.size(); ++i)
                ok = SamePose(r1[i].mRelativePose, r2[i].mRelativePose) &&
                     SamePose(r1[i].mStartNodePose, r2[i].mStartNodePose) &&
                     r1[i].mStartNodeIdx == r2[i].mStartNodeIdx && r1[i].mEndNodeIdx == r2[i].mEndNodeIdx &&
                     SameMat(r1[i].mEstimatedCovMat, r2[i].mEstimatedCovMat);
            std::printf("grid-search matcher + loop detector (%s tail): loops ref %zu / cuda %zu %s; Detect reference %.1f ms, "
                        "cuda %.1f ms\n", round ? "device" : "host", r1.size(), r2.size(), ok ? "IDENTICAL" : "MISMATCH",
                        std::chrono::duration<double, std::milli>(s1 - s0).count(),
                        std::chrono::duration<double, std::milli>(s2 - s1).count());
            failures += !ok;
            if (r1.empty() || r1.size() == 6) { std::printf("expected some but not all nodes to close a loop\n"); ++failures; }
        }
    }
    /* ---- the JSON factories of create_cuda_backends.hpp (instantiated with a flat key table) ---- */
    {
        FlatTree t;
        t.kv = {{"LowResolutionMapWinSize", "5"}, {"SearchRangeX", "1.0"}, {"SearchRangeY", "1.0"},
                {"SearchRangeTheta", "0.6"}, {"ScanRangeMax", "20.0"}, {"ScoreThreshold", "0.6"},
                {"ScanMatcherConfigGroup", "M"}, {"NodeHeightMax", "6"}, {"ScoreConfigGroup", "S"},
                {"CostType", "GreedyEndpoint"}, {"CostConfigGroup", "C"}, {"UsableRangeMin", "0.01"},
                {"UsableRangeMax", "20.0"}, {"Map.NumOfScansForLatestMap", "10"},
                {"ProbabilityHit", "0.6"}, {"ProbabilityMiss", "0.45"}, {"SearchStepX", "0.05"},
                {"SearchStepY", "0.05"}, {"SearchStepTheta", "0.01"}, {"Devices", "0"}};
        auto costFactory = [&](const FlatTree&, const std::string&, const std::string&) { return CostFuncPtr(cost); };
        auto m = LgsB200::CreateScanMatcherRealTimeCorrelativeCuda(t, "M", costFactory);
        auto d1 = LgsB200::CreateLoopDetectorBranchBoundCuda(t, "D", costFactory);
        auto d2 = LgsB200::CreateLoopDetectorRealTimeCorrelativeCuda(t, "D", costFactory);
        auto b = LgsB200::CreateGridMapBuilderCuda(t, "G");
        auto m3 = LgsB200::CreateScanMatcherGridSearchCuda(t, "M", costFactory);
        auto d3 = LgsB200::CreateLoopDetectorGridSearchCuda(t, "D", costFactory);
        bool ok = m && d1 && d2 && b && m3 && d3 && b->LocalMaps().empty();
        /* the factory-made matcher evaluates its tail on the device with the parameters it read from
         * the settings; it must agree with the reference matcher holding the launcher-made cost object */
        {
            ScanMatcherRealTimeCorrelative ref(cost, 5, 1.0, 1.0, 0.6, 20.0);
            const RobotPose2D<double> truth = path[57];
            const auto scan = MakeScan(world, truth, 541, g);
            const RobotPose2D<double> init(truth.mX + 0.11, truth.mY - 0.17, truth.mTheta + 0.05);
            ScanMatchingQuery q1(GridMapType(builder.LatestMap()), scan, init);
            ScanMatchingQuery q2(GridMapType(builder.LatestMap()), scan, init);
            const ScanMatchingSummary a = ref.OptimizePose(q1);
            const ScanMatchingSummary c = m->OptimizePose(q2);
            ok = ok && a.mPoseFound == c.mPoseFound && SameBits(a.mNormalizedCost, c.mNormalizedCost) &&
                 SamePose(a.mEstimatedPose, c.mEstimatedPose) && SameMat(a.mEstimatedCovariance, c.mEstimatedCovariance);
        }
        std::printf("factories: %s\n", ok ? "constructed, factory-made matcher IDENTICAL" : "FAILED");
        failures += !ok;
    }
    std::printf(failures ? "FAILED (%d)\n" : "ALL IDENTICAL\n", failures);
    return failures ? 1 : 0;
}
