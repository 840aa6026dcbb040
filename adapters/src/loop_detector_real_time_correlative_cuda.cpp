/* loop_detector_real_time_correlative_cuda.cpp -- see the header */
#include "lgs_adapters/loop_detector_real_time_correlative_cuda.hpp"

#include <cassert>
#include <vector>

namespace MyLidarGraphSlam {
namespace Mapping {

LoopDetectorRealTimeCorrelativeCuda::LoopDetectorRealTimeCorrelativeCuda(
    const std::shared_ptr<ScanMatcherRealTimeCorrelativeCuda>& scanMatcher,
    const double scoreThreshold) :
    mScanMatcher(scanMatcher),
    mScoreThreshold(scoreThreshold)
{
    assert(scoreThreshold > 0.0);
    assert(scoreThreshold <= 1.0);
}

void LoopDetectorRealTimeCorrelativeCuda::Detect(
    LoopDetectionQueryVector& loopDetectionQueries,
    LoopDetectionResultVector& loopDetectionResults)
{
    loopDetectionResults.clear();

    for (auto& query : loopDetectionQueries) {
        auto& localMapInfo = query.mLocalMapInfo;
        const auto& localMapNode = query.mLocalMapNode;
        assert(localMapNode.Index() >= localMapInfo.mPoseGraphNodeIdxMin &&
               localMapNode.Index() <= localMapInfo.mPoseGraphNodeIdxMax);
        assert(localMapInfo.mFinished);
        /* the coarse map is (re)computed on the device with every batch; nothing to keep on the host */
        localMapInfo.mPrecomputed = true;

        std::vector<Sensor::ScanDataPtr<double>> scans;
        std::vector<RobotPose2D<double>> poses;
        for (const auto& node : query.mPoseGraphNodes) {
            scans.push_back(node.ScanData());
            poses.push_back(node.Pose());
        }
        const std::vector<ScanMatchingSummary> summaries =
            this->mScanMatcher->OptimizePoses(localMapInfo.mMap, scans, poses, this->mScoreThreshold);

        /* one loop closing edge per node whose score exceeds the threshold, in node order
         * (loop_detector_real_time_correlative.cpp:63-90) */
        for (std::size_t k = 0; k < summaries.size(); ++k) {
            if (!summaries[k].mPoseFound)
                continue;
            loopDetectionResults.emplace_back(
                InverseCompound(localMapNode.Pose(), summaries[k].mEstimatedPose), localMapNode.Pose(),
                localMapNode.Index(), query.mPoseGraphNodes[k].Index(), summaries[k].mEstimatedCovariance);
        }
    }
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
