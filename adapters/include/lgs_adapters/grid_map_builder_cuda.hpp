/* grid_map_builder_cuda.hpp
 *
 * Drop-in replacement for MyLidarGraphSlam::Mapping::GridMapBuilder
 * (mapping/grid_map_builder.hpp:111-229): same constructor parameters (plus the device), same public
 * surface (AppendScan, AfterLoopClosure, ConstructGlobalMap, LocalMaps, LocalMapAt, LatestMap,
 * AccumTravelDist, LatestScanIdxMin/Max), same LocalMapInfo bookkeeping -- but every cell update
 * (the loops of UpdateGridMap :170-186 and ConstructMapFromScans :311-328) runs on a B200 through
 * lgs_grid_integrate_scans.  The host GridMapType objects the rest of the system reads are kept
 * bit-identical to what the reference builder would hold, including which patches are allocated:
 *   - map geometry comes from the reference's own GridMap::Expand / Resize on the host map;
 *   - the current local map has a device mirror that follows Expand through lgs_grid_resize, so a
 *     frame costs one scan of integration plus a download of the map;
 *   - a device value is written into a host cell with the cell's own public API: Reset() followed by
 *     Update(v) takes the first-observation path, which stores clamp(v) = v for every value the
 *     filter can produce (binary_bayes_grid_cell.hpp:75-92); cells that were never observed stay
 *     untouched, so patch allocation matches the reference as well.
 * GridMapBuilder is a concrete class without virtual functions, so the launcher selects this one
 * where it creates the builder (slam_launcher.cpp:711-737): see INTEGRATION.md. */
#ifndef LGS_ADAPTERS_GRID_MAP_BUILDER_CUDA_HPP
#define LGS_ADAPTERS_GRID_MAP_BUILDER_CUDA_HPP

#include <memory>
#include <vector>

#include "lgs_b200.h"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

class GridMapBuilderCuda final
{
public:
    GridMapBuilderCuda(double mapResolution,
                       int patchSize,
                       int numOfScansForLatestMap,
                       double travelDistThreshold,
                       double usableRangeMin,
                       double usableRangeMax,
                       double probHit,
                       double probMiss,
                       int device = 0);
    ~GridMapBuilderCuda();
    GridMapBuilderCuda(const GridMapBuilderCuda&) = delete;
    GridMapBuilderCuda& operator=(const GridMapBuilderCuda&) = delete;

    /* Append the new scan data; returns whether a new local map was created */
    bool AppendScan(const std::shared_ptr<PoseGraph>& poseGraph);
    /* Re-create the local grid maps and latest map after the loop closure */
    void AfterLoopClosure(const std::shared_ptr<PoseGraph>& poseGraph);
    /* Construct the global map */
    GridMapType ConstructGlobalMap(const std::shared_ptr<PoseGraph>& poseGraph);

    /* The host maps.  With lazy host maps (SetLazyHostMaps) these accessors first bring the host copies
     * up to date with the device (FlushHostMaps); what they return is always identical to the CPU builder's */
    inline const std::vector<LocalMapInfo>& LocalMaps() const { this->FlushHostMaps(); return this->mLocalMaps; }
    inline LocalMapInfo& LocalMapAt(int localMapIdx) { this->FlushHostMaps(); return this->mLocalMaps.at(localMapIdx); }
    inline const LocalMapInfo& LocalMapAt(int localMapIdx) const
    { this->FlushHostMaps(); return this->mLocalMaps.at(localMapIdx); }
    inline const GridMapType& LatestMap() const { this->FlushHostMaps(); return this->mLatestMap; }

    /* Lazy host maps: AppendScan only integrates on the device and remembers which cells of the host maps
     * are stale; they are downloaded and written back when a host reader asks (the accessors above, a new
     * local map, ConstructGlobalMap) instead of after every frame.  A front end that hands the device map
     * to the matcher (DeviceLatestMap) reads the host maps only when the loop detector runs. */
    void SetLazyHostMaps(bool lazy) { if (!lazy) this->FlushHostMaps(); this->mLazy = lazy; }
    void FlushHostMaps() const;
    inline double AccumTravelDist() const { return this->mAccumTravelDist; }
    inline int LatestScanIdxMin() const { return this->mLatestScanIdxMin; }
    inline int LatestScanIdxMax() const { return this->mLatestScanIdxMax; }

    /* The latest map as it sits on the device after AppendScan / AfterLoopClosure (bit-identical
     * to LatestMap()), for ScanMatcherRealTimeCorrelativeCuda::OptimizePose(const lgs_grid*, ...);
     * nullptr while the device copy holds something else (after ConstructGlobalMap, until the next
     * AppendScan).  Valid until the next call that modifies the builder. */
    const lgs_grid* DeviceLatestMap() const
    { return this->mScratchIsLatest ? this->mDevScratch : nullptr; }

    /* Wall-clock split of the work so far, milliseconds: {host staging of hit points + geometry,
     * device integration (incl. its synchronisation), map download, write-back into the host maps} */
    const double* TimingsMs() const { return this->mTimingsMs; }

    /* Cell updates applied on the device so far (= BinaryBayesGridCell::Update calls of the CPU) */
    long long NumOfCellUpdates() const { return this->mNumOfUpdates; }

private:
    /* Hit points of one pose graph node in the layout of lgs_hit_batch; grows `bbox` */
    void AppendNodeHits(const PoseGraph::Node& node, double* bbox);
    bool UpdateGridMap(const std::shared_ptr<PoseGraph>& poseGraph);
    void UpdateLatestMap(const std::shared_ptr<PoseGraph>& poseGraph);
    void ConstructMapFromScans(GridMapType& gridMap, const std::shared_ptr<PoseGraph>& poseGraph,
                               int nodeIdxMin, int nodeIdxMax);
    /* (Re)create `grid` so that it mirrors the geometry of `map`, all cells unknown */
    void MirrorGeometry(lgs_grid*& grid, const GridMapType& map);
    void ReserveDense(std::size_t cells) const;
    /* Integrate the staged hits into `grid` */
    void Integrate(lgs_grid* grid);
    /* Copy the cells [x0, x1] x [y0, y1] of `grid` into `map` (download + patch write-back) */
    void SyncRegion(const lgs_grid* grid, GridMapType& map, int x0, int y0, int x1, int y1) const;

    const double              mResolution;
    const int                 mPatchSize;
    std::vector<LocalMapInfo> mLocalMaps;
    GridMapType               mLatestMap;
    double                    mAccumTravelDist;
    const int                 mNumOfScansForLatestMap;
    int                       mLatestScanIdxMin;
    int                       mLatestScanIdxMax;
    RobotPose2D<double>       mLastRobotPose;
    double                    mTravelDistLastLocalMap;
    RobotPose2D<double>       mRobotPoseLastLocalMap;
    const double              mTravelDistThreshold;
    const double              mUsableRangeMin;
    const double              mUsableRangeMax;
    const double              mProbHit;
    const double              mProbMiss;

    lgs_ctx*                  mCtx;
    lgs_grid*                 mDevLocal;     /* mirror of mLocalMaps.back().mMap */
    lgs_grid*                 mDevScratch;   /* target of ConstructMapFromScans */
    std::vector<double>       mSensorXY;     /* staged lgs_hit_batch */
    std::vector<int>          mHitBegin;
    std::vector<double>       mHitXY;
    mutable std::vector<double> mDense;      /* page-locked download staging */
    mutable bool              mDensePinned;
    bool                      mLazy;              /* host maps are synchronised on demand */
    mutable bool              mLocalPending;      /* cells mLocalBox of the current local map are stale on the host */
    mutable int               mLocalBox[4];       /* x0, y0, x1, y1 */
    mutable bool              mLatestPending;     /* the host latest map is stale (all of it) */
    bool                      mScratchIsLatest;   /* mDevScratch == mLatestMap */
    long long                 mNumOfUpdates;
    mutable double            mTimingsMs[4];
};

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */

#endif
