// tbv_reg.cuh — device-side registration interface shared by k_register.cu and k_odom.cu.
#pragma once
#include "tbv_cells.cuh"

namespace tbv {

constexpr int GRID_CAP = 16384;  // buckets of the 4 m search grid over a fixed scan's cell means (128 x 128 = 512 m span)

struct CellGrid {       // header of one search grid; ok == 0: no grid (extent / size outside the limits) -> exhaustive search;
                        // 1: entries in bucket order; 2: and ascending in x inside every bucket row (what k_register stages)
  float minx, miny;
  int nx, ny, ok;
};

struct SetView {        // one cell set (MapPointNormal) on the device, field-major
  const double* f;      // f[field*cap + i]
  int cap;
  const int* n_ptr;     // number of cells lives on the device (pipeline) ...
  int n_val;            // ... or is known on the host (API calls); used when n_ptr == nullptr
  // search grid over the cell means (optional; built by cellgrid_build_launch): bucket b holds entries [gstart[b], gstart[b+1])
  CellGrid* grid;
  uint16_t* gstart;     // [GRID_CAP + 1]
  float4* gent;         // [cap] entries in bucket order: (mean x, mean y) narrowed to float (pcl::PointXY), cell index as int bits
};

struct GridStore {      // storage for the grids of n_sets cell sets
  int n_sets = 0, cell_cap = 0;
  DevBuf<CellGrid> hdr;
  DevBuf<uint16_t> start;
  DevBuf<float4> ent;
  int reserve(int n_sets_, int cell_cap_);
  void release() { hdr.release(); start.release(); ent.release(); }
  SetView view(int i, const double* f, int cap, const int* n_ptr, int n_val) const {
    return SetView{f, cap, n_ptr, n_val, hdr.p + i, start.p + (size_t)i * (GRID_CAP + 1), ent.p + (size_t)i * cell_cap};
  }
};

// (re)builds the grids of sets which_dev[0..n_launch) (which_dev == nullptr: sets 0..n_launch-1; entries < 0 are skipped)
// max_extent > 0: bound on |x|, |y| of the cell means (sizes the kernel's shared-memory counters; a larger set simply gets no grid)
// cell_cap: upper bound on the cells of any of those sets (sizes the shared-memory area in which the entries are ordered inside their buckets)
int cellgrid_build_launch(tbv_ctx* ctx, const SetView* sets_dev, const int* which_dev, int n_launch, int n_sets, int cell_cap, double max_extent = 0.0);

struct RegProblem {     // one n_scan_normal_reg::Register call: fixed scans + one moving scan
  int n_fixed;
  int fixed_first;      // first entry of this problem in fixed_set[] / fixed_pose[]
  int src_set;
  int active;           // 0: skip (e.g. first frame of a sequence)
  double src_pose[3];   // initial guess (x, y, theta)
};

struct RegParamsDev {
  int cost, loss, weight_opt;
  double loss_limit, cov_scale, regularization;
  int max_itr_association, max_itr_solver;
  double angle_outlier;   // cos(pi/6) computed on the host (glibc), n_scan_normal.cpp:216
  double radius;          // registration.h:122 radius_ = 2.0
};

struct RegResult {
  double pose[3];         // parameters.back() at exit (x, y, theta)
  double align[3];        // (pose^-1 * fixed_pose[0]) as (x, y, theta): loopclosure.cpp:73 Talign
  int pose_updated;       // 0: the caller's Tsrc stays untouched (no successful solve), n_scan_normal.cpp:118-121
  int success, itrs, lm_iterations, num_residuals, last_n_iterations, termination;
  double score, final_cost, last_relative_decrease;
};

// n_scan_normal.h:75 / n_scan_normal.cpp:9 defaults; cos(pi/6) evaluated by the host libm.
RegParamsDev to_dev(const tbv_reg_params& p);

enum { REG_MODE_REGISTER = 0, REG_MODE_EVAL = 1 };

constexpr int BLK_FIELDS = 9;  // per residual block: src(2) tar(2) [nrm(2) | L00 L10 L11] weight sqrt(weight)

struct RegScratch {            // association / residual-block scratch, [n_problems][...]
  DevBuf<int> assoc;           // [n_problems][max_fixed][slot_cap] target index per (fixed, src) or -1
  DevBuf<double> blocks;       // [n_problems] x (one segment per warp of tiles of 32 residual blocks, BLK_FIELDS x 32 doubles per tile)
  DevBuf<int> n_blocks;        // [n_problems]
  DevBuf<double> residuals;    // [n_problems][2*max_fixed*slot_cap] (eval mode, optional)
  DevBuf<unsigned long long> dbg;  // development builds only (-DTBV_DEV_TIMERS): per-phase cycle counters
  void release() { assoc.release(); blocks.release(); n_blocks.release(); residuals.release(); dbg.release(); }
};

// Launches one CTA per problem.  All pointers are device pointers.  slot_cap >= number of cells of any moving scan,
// tgt_cap >= number of cells of any fixed scan.  In REG_MODE_EVAL the kernel performs a single association at the given
// poses with the search radius selected by eval_itr and one evaluation; results: final_cost = cost, num_residuals, and
// eval_out[p*10 ..] = {cost, g0,g1,g2, H00,H01,H02,H11,H12,H22}; residuals (optional) go to scratch.residuals.
int register_launch(tbv_ctx* ctx, int mode, int eval_itr, const SetView* sets_dev, const RegProblem* problems_dev, const int* fixed_set_dev,
                    const double* fixed_pose_dev, int n_problems, int max_fixed, int slot_cap, int tgt_cap, const RegParamsDev& params,
                    RegResult* results_dev, double* eval_out_dev, bool want_residuals);

RegScratch* reg_scratch(tbv_ctx* ctx);
int reg_scratch_reserve(tbv_ctx* ctx, int n_problems, int max_fixed, int slot_cap, bool want_residuals);

}  // namespace tbv
