// Cross-language check of the .tbvg simple-graph layout (include/tbv_b200.hpp <-> tbv_slam_public_b200/graph_io.py), CPU only:
//   test_graph_io <in.tbvg> <copy.tbvg> <optimised.tbvg>
// parses the file Python wrote, re-serialises it (must be byte-identical), then optimises the graph with CeresLeastSquaresT over the oracle backend and
// writes the graph with the new poses.  Prints one summary line for the Python side to compare.
#include <cstdio>
#include <fstream>
#include <iterator>

#include "oracle_backend.hpp"

static std::string slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  try {
    const std::string bytes = slurp(argv[1]);
    tbv_b200::simple_graph g = tbv_b200::ParseSimpleGraph(bytes);
    const std::string again = tbv_b200::SerializeSimpleGraph(g);
    std::ofstream(argv[2], std::ios::binary) << again;
    size_t n_con = 0, n_cells = 0, n_pts = 0, n_gt = 0, n_quality = 0;
    for (const auto& nc : g) {
      n_con += nc.second.size(); n_cells += nc.first.cloud_normal_.size(); n_pts += nc.first.cloud_peaks_.size() + nc.first.cloud_nopeaks_.size();
      n_gt += nc.first.has_Tgt_;
      for (const auto& c : nc.second) n_quality += c.quality.size();
    }
    std::vector<tbv_b200::Pose3d> nodes;
    std::vector<tbv_b200::Constraint3d> cons;
    tbv_b200::GraphToOptimizerInput(g, nodes, cons);
    tbv_pgo_params par = tbv_b200::default_pgo_params();
    par.loop_scaling = 1.0;
    tbv_b200::CeresLeastSquaresT<OracleBackend> solver(OracleBackend(), nodes, cons, par);
    solver.Solve();
    for (size_t i = 0; i < g.size(); i++) g[i].first.T = nodes[i];
    std::ofstream(argv[3], std::ios::binary) << tbv_b200::SerializeSimpleGraph(g);
    std::printf("nodes %zu constraints %zu optimised %zu cells %zu points %zu gt %zu quality %zu identical %d cost %.17g %.17g iterations %d %s\n", g.size(), n_con,
                cons.size(), n_cells, n_pts, n_gt, n_quality, (int)(again == bytes), solver.summary_.initial_cost, solver.summary_.final_cost,
                solver.summary_.iterations, solver.summary_.termination.c_str());
    // malformed inputs are errors, not crashes
    int caught = 0;
    for (const std::string& bad : {bytes.substr(0, bytes.size() / 2), std::string("22 serialization::archive") + bytes.substr(25), bytes + "x"}) {
      try { tbv_b200::ParseSimpleGraph(bad); } catch (const tbv_b200::Error& e) { caught += e.code == TBV_ERR_INVALID; }
    }
    std::printf("rejected %d\n", caught);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 1;
  }
  return 0;
}
