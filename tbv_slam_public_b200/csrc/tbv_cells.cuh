// tbv_cells.cuh — device-resident cell sets ("MapPointNormal") shared by the cells and registration kernels.
#pragma once
#include "tbv_common.cuh"

namespace tbv {

constexpr int CELL_FIELDS = 16;  // same order as tbv_cell
enum { CF_U0 = 0, CF_U1, CF_C00, CF_C01, CF_C10, CF_C11, CF_SCALE, CF_N0, CF_N1, CF_O0, CF_O1, CF_LMIN, CF_LMAX, CF_SUMI, CF_AVGI, CF_NS };

// A batch of cell sets, field-major: field f of cell i of set b lives at f64[(b*CELL_FIELDS + f)*cap + i].
struct CellStore {
  int batch = 0, cap = 0;
  DevBuf<double> f64;
  DevBuf<int> count;      // [batch] valid cells
  DevBuf<int> n_samples;  // [batch] voxel-grid sample points examined
  int reserve(int batch_, int cap_) {
    batch = batch_;
    cap = cap_;
    int rc;
    if ((rc = f64.reserve((size_t)batch_ * CELL_FIELDS * cap_))) return rc;
    if ((rc = count.reserve(batch_))) return rc;
    return n_samples.reserve(batch_);
  }
  void release() { f64.release(); count.release(); n_samples.release(); }
  double* set_ptr(int b) const { return f64.p + (size_t)b * CELL_FIELDS * cap; }
};

struct CellsParams {
  float radius;
  double downsample_factor;
  int weight_intensity;
  double origin[2];
  double max_extent;  // bound on |x|,|y| used to size the voxel grid scratch (metres); <=0: 400 m
  int max_samples;    // voxel-grid sample points examined per scan; <=0: same as the cell capacity
};

// Build cells for `batch` clouds resident on the device (cloud b = entries [b*cap_pts, b*cap_pts + count[b])).
// Exactly one of inten_u8 / inten_f32 is non-null.  Results go to `out` (reserved by the callee).
int cells_build_dev(tbv_ctx* ctx, const float* x, const float* y, const uint8_t* inten_u8, const float* inten_f32, const int* count_dev,
                    int cap_pts, int batch, const CellsParams& par, int cell_cap, CellStore& out);

// host AoS (tbv_cell) <-> device field-major set
const int* cells_err_dev(tbv_ctx* ctx);  // [batch] TBV_OK / TBV_ERR_CAPACITY of the last cells_build_dev
int cells_upload(tbv_ctx* ctx, const tbv_cell* cells, int n, double* set_dev, int cap);
int cells_download(tbv_ctx* ctx, const double* set_dev, int cap, int n, tbv_cell* cells);

}  // namespace tbv
