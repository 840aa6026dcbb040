/* create_cuda_backends.hpp -- factory functions for the B200 backends, written the way
 * slam_launcher.cpp writes its own (CreateScanMatcherRealTimeCorrelative :296-323,
 * CreateLoopDetectorBranchBound, CreateScanMatcher :325-342, CreateLoopDetector :482-497).
 *
 * They are templates over the property-tree type so that this header also compiles where
 * Boost is absent; in slam_launcher.cpp instantiate them with boost::property_tree::ptree and
 * pass the launcher's own CreateCostFunction (see INTEGRATION.md for the five-line patch).
 *
 * New settings keys (all optional): "Device" (int, default 0) and "DeviceCost" (bool, default true)
 * in the matcher / detector group; in the branch-and-bound loop detector group also
 * "Devices": [0, 1, ..., 7] -- the GPUs of the box the local maps are spread over (local map i on
 * device i mod G).  "Devices" wins over "Device".
 */
#ifndef LGS_ADAPTERS_CREATE_CUDA_BACKENDS_HPP
#define LGS_ADAPTERS_CREATE_CUDA_BACKENDS_HPP

#include <memory>
#include <string>
#include <vector>

#include "lgs_adapters/grid_map_builder_cuda.hpp"
#include "lgs_adapters/grid_search_cuda.hpp"
#include "lgs_adapters/loop_detector_branch_bound_cuda.hpp"
#include "lgs_adapters/loop_detector_real_time_correlative_cuda.hpp"
#include "lgs_adapters/scan_matcher_real_time_correlative_cuda.hpp"

namespace LgsB200 {

namespace Mapping = MyLidarGraphSlam::Mapping;

/* The CostGreedyEndpoint constructor arguments exactly as CreateCostGreedyEndpoint passes them
 * (slam_launcher.cpp:60-72: same keys, same defaults, and StandardDeviation / ScalingFactor in the
 * positions the launcher puts them, which the constructor reads as scaling factor / standard
 * deviation), for the adapters that evaluate the cost function on the device */
template <typename Ptree>
lgs_cost_params ReadCostGreedyEndpointParams(const Ptree& jsonSettings, const std::string& configGroup)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    lgs_cost_params p;
    p.usable_range_min = config.get("UsableRangeMin", 0.01);
    p.usable_range_max = config.get("UsableRangeMax", 50.0);
    p.hit_and_missed_dist = config.get("HitAndMissedDist", 0.075);
    p.occupancy_threshold = config.get("OccupancyThreshold", 0.1);
    p.kernel_size = config.get("KernelSize", 1);
    p.scaling_factor = config.get("StandardDeviation", 0.05);      /* constructor's 6th argument */
    p.standard_deviation = config.get("ScalingFactor", 1.0);       /* constructor's 7th argument */
    return p;
}

/* "Devices": [0, 1, ...] of a detector group, or { "Device" } when the key is absent.  A JSON array is
 * a child whose elements have empty keys (boost::property_tree) */
template <typename Ptree>
std::vector<int> ReadDevices(const Ptree& config)
{
    std::vector<int> devices;
    if (auto list = config.get_child_optional("Devices"))
        for (const auto& item : *list)
            devices.push_back(item.second.template get_value<int>());
    if (devices.empty())
        devices.push_back(config.get("Device", 0));
    return devices;
}

/* "ScanMatcherType": "RealTimeCorrelativeCuda" -- same keys and defaults as
 * launcher_settings_default.json:42-50 / slam_launcher.cpp:302-306 */
template <typename Ptree, typename CostFactory>
std::shared_ptr<Mapping::ScanMatcher> CreateScanMatcherRealTimeCorrelativeCuda(
    const Ptree& jsonSettings, const std::string& configGroup, CostFactory createCostFunction)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    const int lowResolution = config.get("LowResolutionMapWinSize", 10);
    const double rangeX = config.get("SearchRangeX", 0.75);
    const double rangeY = config.get("SearchRangeY", 0.75);
    const double rangeTheta = config.get("SearchRangeTheta", 0.5);
    const double scanRangeMax = config.get("ScanRangeMax", 20.0);
    const int device = config.get("Device", 0);
    const std::string costType = config.get("CostType", std::string("GreedyEndpoint"));
    const std::string costConfigGroup =
        config.get("CostConfigGroup", std::string("CostGreedyEndpoint"));
    auto pCostFunc = createCostFunction(jsonSettings, costType, costConfigGroup);
    auto pMatcher = std::make_shared<Mapping::ScanMatcherRealTimeCorrelativeCuda>(
        pCostFunc, lowResolution, rangeX, rangeY, rangeTheta, scanRangeMax, device);
    /* "DeviceCost" (default true): evaluate the tail (Cost / ComputeCovariance) on the device */
    if (costType == "GreedyEndpoint" && config.get("DeviceCost", true))
        pMatcher->UseDeviceCost(ReadCostGreedyEndpointParams(jsonSettings, costConfigGroup));
    return pMatcher;
}

/* "LoopDetectorType": "BranchBoundCuda" -- reads the detector group
 * (launcher_settings_default.json:128-132), its scan matcher group (:134-150) and the
 * pixel-accurate score group (:152-160) exactly like the CPU factories do */
template <typename Ptree, typename CostFactory>
std::shared_ptr<Mapping::LoopDetector> CreateLoopDetectorBranchBoundCuda(
    const Ptree& jsonSettings, const std::string& configGroup, CostFactory createCostFunction)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    const double scoreThreshold = config.template get<double>("ScoreThreshold");
    const std::vector<int> devices = ReadDevices(config);
    const std::string matcherGroup = config.template get<std::string>("ScanMatcherConfigGroup");

    const Ptree& matcher = jsonSettings.get_child(matcherGroup);
    const int nodeHeightMax = matcher.template get<int>("NodeHeightMax");
    const double rangeX = matcher.template get<double>("SearchRangeX");
    const double rangeY = matcher.template get<double>("SearchRangeY");
    const double rangeTheta = matcher.template get<double>("SearchRangeTheta");
    const double scanRangeMax = matcher.template get<double>("ScanRangeMax");
    const std::string scoreGroup = matcher.template get<std::string>("ScoreConfigGroup");
    const std::string costType = matcher.template get<std::string>("CostType");
    const std::string costGroup = matcher.template get<std::string>("CostConfigGroup");

    const Ptree& score = jsonSettings.get_child(scoreGroup);
    const double usableRangeMin = score.template get<double>("UsableRangeMin");
    const double usableRangeMax = score.template get<double>("UsableRangeMax");

    auto pCostFunc = createCostFunction(jsonSettings, costType, costGroup);
    auto pDetector = std::make_shared<Mapping::LoopDetectorBranchBoundCuda>(
        usableRangeMin, usableRangeMax, pCostFunc, nodeHeightMax, rangeX, rangeY, rangeTheta,
        scanRangeMax, scoreThreshold, devices);
    if (costType == "GreedyEndpoint" && config.get("DeviceCost", true))
        pDetector->UseDeviceCost(ReadCostGreedyEndpointParams(jsonSettings, costGroup));
    return pDetector;
}

/* "LoopDetectorType": "RealTimeCorrelativeCuda" -- the detector group names its matcher group
 * exactly like CreateLoopDetectorRealTimeCorrelative does (slam_launcher.cpp:418-446); the matcher is
 * created by CreateScanMatcherRealTimeCorrelativeCuda above */
template <typename Ptree, typename CostFactory>
std::shared_ptr<Mapping::LoopDetector> CreateLoopDetectorRealTimeCorrelativeCuda(
    const Ptree& jsonSettings, const std::string& configGroup, CostFactory createCostFunction)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    const double scoreThreshold = config.template get<double>("ScoreThreshold");
    const std::string matcherGroup = config.template get<std::string>("ScanMatcherConfigGroup");
    auto pScanMatcher = std::dynamic_pointer_cast<Mapping::ScanMatcherRealTimeCorrelativeCuda>(
        CreateScanMatcherRealTimeCorrelativeCuda(jsonSettings, matcherGroup, createCostFunction));
    return std::make_shared<Mapping::LoopDetectorRealTimeCorrelativeCuda>(pScanMatcher, scoreThreshold);
}

/* "ScanMatcherType": "GridSearchCuda" -- the keys CreateScanMatcherGridSearch reads
 * (slam_launcher.cpp:185-228, launcher_settings_default.json:71-82) plus the pixel-accurate score
 * group's usable range (:83-86) */
template <typename Ptree, typename CostFactory>
std::shared_ptr<Mapping::ScanMatcher> CreateScanMatcherGridSearchCuda(
    const Ptree& jsonSettings, const std::string& configGroup, CostFactory createCostFunction)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    const double rangeX = config.template get<double>("SearchRangeX");
    const double rangeY = config.template get<double>("SearchRangeY");
    const double rangeTheta = config.template get<double>("SearchRangeTheta");
    const double stepX = config.template get<double>("SearchStepX");
    const double stepY = config.template get<double>("SearchStepY");
    const double stepTheta = config.template get<double>("SearchStepTheta");
    const int device = config.get("Device", 0);
    const std::string scoreGroup = config.template get<std::string>("ScoreConfigGroup");
    const std::string costType = config.template get<std::string>("CostType");
    const std::string costGroup = config.template get<std::string>("CostConfigGroup");

    const Ptree& score = jsonSettings.get_child(scoreGroup);
    const double usableRangeMin = score.template get<double>("UsableRangeMin");
    const double usableRangeMax = score.template get<double>("UsableRangeMax");

    auto pCostFunc = createCostFunction(jsonSettings, costType, costGroup);
    auto pMatcher = std::make_shared<Mapping::ScanMatcherGridSearchCuda>(
        usableRangeMin, usableRangeMax, pCostFunc, rangeX, rangeY, rangeTheta, stepX, stepY, stepTheta,
        device);
    if (costType == "GreedyEndpoint" && config.get("DeviceCost", true))
        pMatcher->UseDeviceCost(ReadCostGreedyEndpointParams(jsonSettings, costGroup));
    return pMatcher;
}

/* "LoopDetectorType": "GridSearchCuda" -- like CreateLoopDetectorGridSearch (slam_launcher.cpp:386-415) */
template <typename Ptree, typename CostFactory>
std::shared_ptr<Mapping::LoopDetector> CreateLoopDetectorGridSearchCuda(
    const Ptree& jsonSettings, const std::string& configGroup, CostFactory createCostFunction)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    const double scoreThreshold = config.template get<double>("ScoreThreshold");
    const std::string matcherGroup = config.template get<std::string>("ScanMatcherConfigGroup");
    auto pScanMatcher = std::dynamic_pointer_cast<Mapping::ScanMatcherGridSearchCuda>(
        CreateScanMatcherGridSearchCuda(jsonSettings, matcherGroup, createCostFunction));
    return std::make_shared<Mapping::LoopDetectorGridSearchCuda>(pScanMatcher, scoreThreshold);
}

/* "GridMapBuilder": { ..., "Backend": "Cuda", "Device": 0 } -- the other keys are the ones
 * CreateGridMapBuilder reads (slam_launcher.cpp:711-737, launcher_settings_default.json:175-185).
 * GridMapBuilder is a concrete class, so the caller holds the result by its own type (or behind
 * the two-line interface sketched in INTEGRATION.md). */
template <typename Ptree>
std::shared_ptr<Mapping::GridMapBuilderCuda> CreateGridMapBuilderCuda(
    const Ptree& jsonSettings, const std::string& configGroup)
{
    const Ptree& config = jsonSettings.get_child(configGroup);
    /* same keys and defaults as slam_launcher.cpp:718-729 */
    const double mapResolution = config.get("Map.Resolution", 0.05);
    const int patchSize = config.get("Map.PatchSize", 64);
    const int numOfLatestScans = config.get("Map.NumOfScansForLatestMap", 5);
    const double travelDistThreshold = config.get("Map.TravelDistThresholdForLocalMap", 20.0);
    const double usableRangeMin = config.get("UsableRangeMin", 0.01);
    const double usableRangeMax = config.get("UsableRangeMax", 50.0);
    const double probHit = config.get("ProbabilityHit", 0.9);
    const double probMiss = config.get("ProbabilityMiss", 0.1);
    const int device = config.get("Device", 0);
    return std::make_shared<Mapping::GridMapBuilderCuda>(
        mapResolution, patchSize, numOfLatestScans, travelDistThreshold, usableRangeMin,
        usableRangeMax, probHit, probMiss, device);
}

} /* namespace LgsB200 */

#endif
