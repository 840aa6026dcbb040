/* lgs_slam_launch.cpp -- `lgs_slam_launch <CARMEN log> <settings.json> [output] [--set Key=Value ...]`
 *
 * The reference's slam_launch (slam_launcher.cpp:927-1026) for trees without Boost / libpng / gnuplot:
 * it loads the SAME settings file (launcher_settings_default.json) with the same keys and defaults, reads
 * the log with the reference's own CarmenLogReader (io/carmen/carmen_reader.cpp, compiled unmodified),
 * builds the same object graph -- LidarGraphSlam { Frontend { ScanInterpolator, ScanMatcher },
 * Backend { PoseGraphOptimizer, LoopSearcher, LoopDetector }, GridMapBuilder, PoseGraph } out of the
 * reference's own classes -- and runs the same ProcessScan loop.  What it adds is the type strings of the
 * B200 backends (lgs_adapters/create_cuda_backends.hpp):
 *     Frontend.LocalSlam.ScanMatcherType = "RealTimeCorrelativeCuda" | "GridSearchCuda"
 *     Backend.LoopDetectorType           = "BranchBoundCuda" | "RealTimeCorrelativeCuda" | "GridSearchCuda"
 * so config C1 (default settings on a 180-beam log) runs end to end from a log file and a settings
 * file, with the correlative front end on the CPU (default JSON unchanged) or on the GPU (one key).
 *
 * Back end: by default its iterations (loop search -> Detect -> AppendLoopClosingEdges -> Optimize ->
 * AfterLoopClosure, lidar_graph_slam_backend.cpp:21-61) run on the main thread at the frames where the
 * front end notifies it (lidar_graph_slam_frontend.cpp:127-130), so a run is reproducible and two runs
 * with different matcher / detector types can be compared pose by pose; AfterLoopClosure is applied once
 * the next frame has been appended, because it asserts that a newer odometry edge exists
 * (lidar_graph_slam.cpp:341-347 -- in slam_launch the optimiser is slow enough for that to hold).
 * `--async-backend` runs the reference's own back-end thread instead (StartBackend / StopBackend).
 *
 * Differences from slam_launch, all stated on stderr when they apply:
 *   - no gnuplot GUI and no PNG output: the result is written as text, <output>.poses.txt (pose-graph
 *     nodes, 17 significant digits) and <output>.edges.txt;
 *   - "PoseGraphOptimizerType": "LM" needs Eigen's sparse solvers (pose_graph_optimizer_lm.cpp:68-222);
 *     where Eigen is absent only "None" is available (loop-closing edges are still detected and appended,
 *     the graph is not re-optimised).  Nothing falls back silently: an unavailable type is an error;
 *   - the scan matchers that need Eigen beyond 3x3 matrices (LinearSolver, CostSquareError) are not built.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "mini_ptree.hpp"

#include "lgs_adapters/create_cuda_backends.hpp"

#include "my_lidar_graph_slam/io/carmen/carmen_reader.hpp"
#include "my_lidar_graph_slam/mapping/cost_function_greedy_endpoint.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/lidar_graph_slam.hpp"
#include "my_lidar_graph_slam/mapping/lidar_graph_slam_backend.hpp"
#include "my_lidar_graph_slam/mapping/lidar_graph_slam_frontend.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_empty.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_real_time_correlative.hpp"
#include "my_lidar_graph_slam/mapping/loop_searcher_nearest.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph_optimizer.hpp"
#include "my_lidar_graph_slam/mapping/scan_accumulator.hpp"
#include "my_lidar_graph_slam/mapping/scan_interpolator.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_hill_climbing.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_real_time_correlative.hpp"
#include "my_lidar_graph_slam/mapping/score_function_pixel_accurate.hpp"

using namespace MyLidarGraphSlam;
using LgsLauncher::Ptree;

namespace {

[[noreturn]] void Fail(const std::string& message)
{
    std::cerr << "lgs_slam_launch: " << message << std::endl;
    std::exit(EXIT_FAILURE);
}

/* "None": the pose graph keeps the poses the front end gave it.  For builds without Eigen's sparse
 * solvers; ComputeErrorFunction is the residual of one edge in the start node's frame */
class PoseGraphOptimizerNone final : public Mapping::PoseGraphOptimizer
{
public:
    void Optimize(std::vector<Mapping::PoseGraph::Node>&,
                  const std::vector<Mapping::PoseGraph::Edge>&) override { }
    void ComputeErrorFunction(const RobotPose2D<double>& startNodePose,
                              const RobotPose2D<double>& endNodePose,
                              const RobotPose2D<double>& edgeRelPose,
                              Eigen::Vector3d& errorVec) const override
    {
        const RobotPose2D<double> rel = InverseCompound(startNodePose, endNodePose);
        errorVec = Eigen::Vector3d(edgeRelPose.mX - rel.mX, edgeRelPose.mY - rel.mY,
                                   NormalizeAngle(edgeRelPose.mTheta - rel.mTheta));
    }
};

/* ---- factories for the reference's own classes: keys and defaults of slam_launcher.cpp ---- */

Mapping::CostFuncPtr CreateCostFunction(const Ptree& settings, const std::string& type,
                                        const std::string& group)
{
    if (type != "GreedyEndpoint")
        Fail("CostType \"" + type + "\" is not built here (CostSquareError needs Eigen beyond 3x3 matrices)");
    const Ptree& config = settings.get_child(group);
    /* slam_launcher.cpp:60-72 passes StandardDeviation and ScalingFactor in this order */
    return std::make_shared<Mapping::CostGreedyEndpoint>(
        config.get("UsableRangeMin", 0.01), config.get("UsableRangeMax", 50.0),
        config.get("HitAndMissedDist", 0.075), config.get("OccupancyThreshold", 0.1),
        config.get("KernelSize", 1), config.get("StandardDeviation", 0.05),
        config.get("ScalingFactor", 1.0));
}

std::shared_ptr<Mapping::ScorePixelAccurate> CreateScore(const Ptree& settings, const std::string& type,
                                                          const std::string& group)
{
    if (type != "PixelAccurate")
        Fail("ScoreType \"" + type + "\" is unknown");
    const Ptree& config = settings.get_child(group);
    return std::make_shared<Mapping::ScorePixelAccurate>(
        config.get<double>("UsableRangeMin"), config.get<double>("UsableRangeMax"));
}

std::shared_ptr<Mapping::ScanMatcherRealTimeCorrelative> CreateRealTimeCorrelative(
    const Ptree& settings, const std::string& group)
{
    const Ptree& config = settings.get_child(group);
    auto cost = CreateCostFunction(settings, config.get("CostType", "GreedyEndpoint"),
                                   config.get("CostConfigGroup", "CostGreedyEndpoint"));
    return std::make_shared<Mapping::ScanMatcherRealTimeCorrelative>(
        cost, config.get("LowResolutionMapWinSize", 10), config.get("SearchRangeX", 0.75),
        config.get("SearchRangeY", 0.75), config.get("SearchRangeTheta", 0.5),
        config.get("ScanRangeMax", 20.0));
}

std::shared_ptr<Mapping::ScanMatcherGridSearch> CreateGridSearch(const Ptree& settings, const std::string& group)
{
    const Ptree& config = settings.get_child(group);
    auto score = CreateScore(settings, config.get<std::string>("ScoreType"), config.get<std::string>("ScoreConfigGroup"));
    auto cost = CreateCostFunction(settings, config.get<std::string>("CostType"), config.get<std::string>("CostConfigGroup"));
    return std::make_shared<Mapping::ScanMatcherGridSearch>(
        score, cost, config.get<double>("SearchRangeX"), config.get<double>("SearchRangeY"),
        config.get<double>("SearchRangeTheta"), config.get<double>("SearchStepX"),
        config.get<double>("SearchStepY"), config.get<double>("SearchStepTheta"));
}

std::shared_ptr<Mapping::ScanMatcherBranchBound> CreateBranchBound(const Ptree& settings, const std::string& group)
{
    const Ptree& config = settings.get_child(group);
    auto score = CreateScore(settings, config.get<std::string>("ScoreType"), config.get<std::string>("ScoreConfigGroup"));
    auto cost = CreateCostFunction(settings, config.get<std::string>("CostType"), config.get<std::string>("CostConfigGroup"));
    return std::make_shared<Mapping::ScanMatcherBranchBound>(
        score, cost, config.get<int>("NodeHeightMax"), config.get<double>("SearchRangeX"),
        config.get<double>("SearchRangeY"), config.get<double>("SearchRangeTheta"),
        config.get<double>("ScanRangeMax"));
}

std::shared_ptr<Mapping::ScanMatcher> CreateScanMatcher(const Ptree& settings, const std::string& type,
                                                        const std::string& group)
{
    auto costFactory = [](const Ptree& s, const std::string& t, const std::string& g) { return CreateCostFunction(s, t, g); };
    if (type == "RealTimeCorrelative")
        return CreateRealTimeCorrelative(settings, group);
    if (type == "RealTimeCorrelativeCuda")
        return LgsB200::CreateScanMatcherRealTimeCorrelativeCuda(settings, group, costFactory);
    if (type == "GridSearch")
        return CreateGridSearch(settings, group);
    if (type == "GridSearchCuda")
        return LgsB200::CreateScanMatcherGridSearchCuda(settings, group, costFactory);
    if (type == "BranchBound")
        return CreateBranchBound(settings, group);
    if (type == "HillClimbing") {
        const Ptree& config = settings.get_child(group);
        auto cost = CreateCostFunction(settings, config.get("CostType", "GreedyEndpoint"),
                                       config.get("CostConfigGroup", "CostGreedyEndpoint"));
        return std::make_shared<Mapping::ScanMatcherHillClimbing>(
            config.get("LinearStep", 0.1), config.get("AngularStep", 0.1), config.get("MaxIterations", 100),
            config.get("MaxNumOfRefinements", 5), cost);
    }
    Fail("ScanMatcherType \"" + type + "\" is not built here");
}

std::shared_ptr<Mapping::LoopDetector> CreateLoopDetector(const Ptree& settings, const std::string& type,
                                                          const std::string& group)
{
    auto costFactory = [](const Ptree& s, const std::string& t, const std::string& g) { return CreateCostFunction(s, t, g); };
    if (type == "Empty")
        return std::make_shared<Mapping::LoopDetectorEmpty>();
    const Ptree& config = settings.get_child(group);
    const std::string matcherGroup = config.get<std::string>("ScanMatcherConfigGroup");
    if (type == "BranchBound")
        return std::make_shared<Mapping::LoopDetectorBranchBound>(
            CreateBranchBound(settings, matcherGroup), config.get<double>("ScoreThreshold"));
    if (type == "BranchBoundCuda")
        return LgsB200::CreateLoopDetectorBranchBoundCuda(settings, group, costFactory);
    if (type == "RealTimeCorrelative")
        return std::make_shared<Mapping::LoopDetectorRealTimeCorrelative>(
            CreateRealTimeCorrelative(settings, matcherGroup), config.get<double>("ScoreThreshold"));
    if (type == "RealTimeCorrelativeCuda")
        return LgsB200::CreateLoopDetectorRealTimeCorrelativeCuda(settings, group, costFactory);
    if (type == "GridSearch")
        return std::make_shared<Mapping::LoopDetectorGridSearch>(
            CreateGridSearch(settings, matcherGroup), config.get<double>("ScoreThreshold"));
    if (type == "GridSearchCuda")
        return LgsB200::CreateLoopDetectorGridSearchCuda(settings, group, costFactory);
    Fail("LoopDetectorType \"" + type + "\" is unknown");
}

std::shared_ptr<Mapping::PoseGraphOptimizer> CreateOptimizer(const std::string& type)
{
    if (type == "None")
        return std::make_shared<PoseGraphOptimizerNone>();
    Fail("PoseGraphOptimizerType \"" + type + "\" needs Eigen's sparse solvers, which this build does not have; "
         "pass --set Backend.PoseGraphOptimizerType=None (loop-closing edges are still detected and appended)");
}

/* The back-end parts, kept by the launcher as well so that it can run an iteration synchronously */
struct BackendParts
{
    std::shared_ptr<Mapping::PoseGraphOptimizer> mOptimizer;
    std::shared_ptr<Mapping::LoopSearcher>       mSearcher;
    std::shared_ptr<Mapping::LoopDetector>       mDetector;
    int                                          mLoopDetectionInterval;
};

std::shared_ptr<Mapping::LidarGraphSlam> CreateLidarGraphSlam(const Ptree& settings, BackendParts& parts)
{
    const Ptree& top = settings.get_child("LidarGraphSlam");
    /* grid map builder (slam_launcher.cpp:711-737) */
    const Ptree& mapConfig = settings.get_child(top.get("GridMapBuilderConfigGroup", "GridMapBuilder"));
    auto builder = std::make_shared<Mapping::GridMapBuilder>(
        mapConfig.get("Map.Resolution", 0.05), mapConfig.get("Map.PatchSize", 64),
        mapConfig.get("Map.NumOfScansForLatestMap", 5), mapConfig.get("Map.TravelDistThresholdForLocalMap", 20.0),
        mapConfig.get("UsableRangeMin", 0.01), mapConfig.get("UsableRangeMax", 50.0),
        mapConfig.get("ProbabilityHit", 0.9), mapConfig.get("ProbabilityMiss", 0.1));
    auto poseGraph = std::make_shared<Mapping::PoseGraph>();
    /* front end (slam_launcher.cpp:740-806) */
    const Ptree& front = settings.get_child(top.get("FrontendConfigGroup", "Frontend"));
    std::shared_ptr<Mapping::ScanAccumulator> accumulator;
    if (front.get("UseScanAccumulator", false))
        accumulator = std::make_shared<Mapping::ScanAccumulator>(static_cast<std::size_t>(
            settings.get_child(front.get("ScanAccumulatorConfigGroup", "ScanAccumulator")).get("NumOfAccumulatedScans", 3)));
    std::shared_ptr<Mapping::ScanInterpolator> interpolator;
    if (front.get("UseScanInterpolator", true)) {
        const Ptree& c = settings.get_child(front.get("ScanInterpolatorConfigGroup", "ScanInterpolator"));
        interpolator = std::make_shared<Mapping::ScanInterpolator>(c.get("DistScans", 0.05), c.get("DistThresholdEmpty", 0.25));
    }
    auto matcher = CreateScanMatcher(settings, front.get("LocalSlam.ScanMatcherType", "HillClimbing"),
                                     front.get("LocalSlam.ScanMatcherConfigGroup", "ScanMatcherHillClimbing"));
    const RobotPose2D<double> initialPose { front.get("InitialPose.X", 0.0), front.get("InitialPose.Y", 0.0),
                                            front.get("InitialPose.Theta", 0.0) };
    auto frontend = std::make_shared<Mapping::LidarGraphSlamFrontend>(
        accumulator, interpolator, matcher, initialPose, front.get("UpdateThresholdTravelDist", 1.0),
        front.get("UpdateThresholdAngle", 0.5), front.get("UpdateThresholdTime", 5.0),
        front.get("LoopDetectionInterval", 10));
    /* back end (slam_launcher.cpp:809-846) */
    const Ptree& back = settings.get_child(top.get("BackendConfigGroup", "Backend"));
    auto optimizer = CreateOptimizer(back.get("PoseGraphOptimizerType", "LM"));
    if (back.get("LoopSearcherType", "Nearest") != "Nearest")
        Fail("LoopSearcherType \"" + back.get("LoopSearcherType", "Nearest") + "\" is unknown");
    const Ptree& searcherConfig = settings.get_child(back.get("LoopSearcherConfigGroup", "LoopSearcherNearest"));
    auto searcher = std::make_shared<Mapping::LoopSearcherNearest>(
        searcherConfig.get("TravelDistThreshold", 10.0), searcherConfig.get("PoseGraphNodeDistMax", 2.0),
        searcherConfig.get("NumOfCandidateNodes", 2));
    auto detector = CreateLoopDetector(settings, back.get("LoopDetectorType", "GridSearch"),
                                       back.get("LoopDetectorConfigGroup", "LoopDetectorGridSearch"));
    auto backend = std::make_shared<Mapping::LidarGraphSlamBackend>(optimizer, searcher, detector);
    parts = BackendParts { optimizer, searcher, detector, front.get("LoopDetectionInterval", 10) };
    return std::make_shared<Mapping::LidarGraphSlam>(frontend, backend, builder, poseGraph);
}

} /* namespace */

int main(int argc, char** argv)
{
    std::vector<std::string> positional;
    std::vector<std::pair<std::string, std::string>> overrides;
    bool asyncBackend = false;
    for (int k = 1; k < argc; ++k) {
        const std::string arg = argv[k];
        if (arg == "--async-backend") {
            asyncBackend = true;
        } else if (arg == "--set" && k + 1 < argc) {
            const std::string kv = argv[++k];
            const std::size_t eq = kv.find('=');
            if (eq == std::string::npos)
                Fail("--set expects Key=Value, got " + kv);
            overrides.emplace_back(kv.substr(0, eq), kv.substr(eq + 1));
        } else {
            positional.push_back(arg);
        }
    }
    if (positional.size() < 2) {
        std::cerr << "Usage: " << argv[0] << " <Carmen log file name> <JSON settings file name> [output name] "
                     "[--set Key=Value ...] [--async-backend]" << std::endl;
        return EXIT_FAILURE;
    }
    const std::string output = positional.size() > 2 ? positional[2] : std::string("lgs_slam_launch_out");

    /* Carmen log (slam_launcher.cpp:876-892) */
    std::vector<Sensor::SensorDataPtr> logData;
    {
        std::ifstream logFile(positional[0]);
        if (!logFile)
            Fail("failed to open log file " + positional[0]);
        IO::Carmen::CarmenLogReader reader;
        reader.Load(logFile, logData);
    }
    if (logData.empty())
        Fail("the log holds no sensor data");

    /* settings (slam_launcher.cpp:905-925) */
    Ptree settings;
    try {
        LgsLauncher::ReadJson(positional[1], settings);
    } catch (const std::exception& e) {
        Fail(e.what());
    }
    for (const auto& kv : overrides)
        settings.put(kv.first, kv.second);
    if (settings.get("Launcher.GuiEnabled", true))
        std::cerr << "lgs_slam_launch: Launcher.GuiEnabled is ignored (no gnuplot here)" << std::endl;

    std::shared_ptr<Mapping::LidarGraphSlam> slam;
    BackendParts backendParts;
    try {
        slam = CreateLidarGraphSlam(settings, backendParts);
    } catch (const std::exception& e) {
        Fail(e.what());
    }

    if (asyncBackend)
        slam->StartBackend();
    int numOfScans = 0, numOfFrames = 0, numOfDetects = 0;
    /* synchronous back end: the optimised nodes of an iteration wait here until the next frame exists */
    bool closurePending = false;
    std::vector<Mapping::PoseGraph::Node> closureNodes;
    const auto t0 = std::chrono::steady_clock::now();
    for (const auto& sensorData : logData) {
        auto scanData = std::dynamic_pointer_cast<const Sensor::ScanData<double>>(sensorData);
        if (scanData == nullptr)
            continue;
        ++numOfScans;
        if (!slam->ProcessScan(scanData, scanData->OdomPose()))
            continue;
        ++numOfFrames;
        if (asyncBackend)
            continue;
        if (closurePending) {
            slam->AfterLoopClosure(closureNodes);
            closurePending = false;
        }
        /* the front end's own condition for NotifyBackend, on the count before its increment */
        const int count = slam->ProcessCount() - 1;
        if (count <= backendParts.mLoopDetectionInterval || count % backendParts.mLoopDetectionInterval != 0)
            continue;
        /* one iteration of LidarGraphSlamBackend::Run (lidar_graph_slam_backend.cpp:27-60) */
        auto candidates = backendParts.mSearcher->Search(slam->GetLoopSearchHint());
        auto queries = slam->GetLoopDetectionQueries(candidates);
        Mapping::LoopDetectionResultVector results;
        backendParts.mDetector->Detect(queries, results);
        ++numOfDetects;
        slam->UpdatePrecomputedGridMaps(queries);
        if (results.empty())
            continue;
        slam->AppendLoopClosingEdges(results);
        std::vector<Mapping::PoseGraph::Edge> closureEdges;
        closureNodes.clear();
        slam->GetPoseGraph(closureNodes, closureEdges);
        backendParts.mOptimizer->Optimize(closureNodes, closureEdges);
        closurePending = true;
    }
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (asyncBackend)
        slam->StopBackend();
    else if (closurePending)
        std::cerr << "lgs_slam_launch: the log ended before the last loop closure could be applied" << std::endl;

    /* result as text instead of MapSaver's PNG / JSON (map_saver.cpp:413-535) */
    std::vector<Mapping::PoseGraph::Node> nodes;
    std::vector<Mapping::PoseGraph::Edge> edges;
    slam->GetPoseGraph(nodes, edges);
    {
        std::FILE* f = std::fopen((output + ".poses.txt").c_str(), "w");
        if (f == nullptr)
            Fail("cannot write " + output + ".poses.txt");
        for (const auto& n : nodes)
            std::fprintf(f, "%d %.17g %.17g %.17g\n", n.Index(), n.Pose().mX, n.Pose().mY, n.Pose().mTheta);
        std::fclose(f);
        f = std::fopen((output + ".edges.txt").c_str(), "w");
        if (f == nullptr)
            Fail("cannot write " + output + ".edges.txt");
        for (const auto& e : edges)
            std::fprintf(f, "%d %d %d %.17g %.17g %.17g\n", e.StartNodeIndex(), e.EndNodeIndex(),
                         e.IsOdometricConstraint() ? 0 : 1, e.RelativePose().mX, e.RelativePose().mY,
                         e.RelativePose().mTheta);
        std::fclose(f);
    }
    int latestMin = 0, latestMax = 0;
    const Mapping::GridMapType latest = slam->GetLatestMap(latestMin, latestMax);
    std::size_t loops = 0;
    for (const auto& e : edges) loops += e.IsOdometricConstraint() ? 0 : 1;
    std::printf("{\"scans\": %d, \"frames\": %d, \"seconds\": %.3f, \"frames_per_s\": %.2f, \"nodes\": %zu, "
                "\"edges\": %zu, \"loop_edges\": %zu, \"detect_calls\": %d, \"latest_map_cells\": [%d, %d], "
                "\"scan_matcher\": \"%s\", \"loop_detector\": \"%s\"}\n", numOfScans, numOfFrames, seconds,
                numOfFrames / seconds, nodes.size(), edges.size(), loops, numOfDetects,
                latest.NumOfGridCellsX(), latest.NumOfGridCellsY(),
                settings.get("Frontend.LocalSlam.ScanMatcherType", "HillClimbing").c_str(),
                settings.get("Backend.LoopDetectorType", "GridSearch").c_str());
    return EXIT_SUCCESS;
}
