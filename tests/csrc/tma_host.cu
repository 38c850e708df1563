// Host-side view of the staged kernel's tile geometry (csrc/lbm_launch.cuh) for tests/test_tma_host.py: no GPU needed.
#include "../../lettuce_b200/csrc/lbm_launch.cuh"

extern "C" {
int tma_host_available(int dtype, long long nodes, int n2) { return lbm::tma_available(dtype, nodes, n2) ? 1 : 0; }
int tma_host_row_extent(int n2) { return lbm::tma_row_extent(n2); }
int tma_host_tile_rows(int n2) { return lbm::tma_tile_rows(n2); }
int tma_host_rows_boxable(int n0, int n1, int n2) { return lbm::tma_rows_boxable(n0, n1, n2) ? 1 : 0; }
int tma_host_tile_nodes(void) { return lbm::kTmaTileNodes; }
int tma_host_max_rows(void) { return lbm::kTmaMaxRows; }
int tma_host_stage_bytes(int q, int push) {
    if (q == 9) return (int)sizeof(float) * (push ? lbm::tma_push_stage_floats<lbm::D2Q9>() : lbm::tma_stage_floats<lbm::D2Q9>());
    if (q == 19) return (int)sizeof(float) * (push ? lbm::tma_push_stage_floats<lbm::D3Q19>() : lbm::tma_stage_floats<lbm::D3Q19>());
    return (int)sizeof(float) * (push ? lbm::tma_push_stage_floats<lbm::D3Q27>() : lbm::tma_stage_floats<lbm::D3Q27>());
}
}

// the run-time velocity table of the producer threads restates the compile-time stencils
namespace {
constexpr signed char kTable[3][27][3] = {
    {{0, 0, 0}, {1, 0, 0}, {0, 0, 1}, {-1, 0, 0}, {0, 0, -1}, {1, 0, 1}, {-1, 0, 1}, {-1, 0, -1}, {1, 0, -1}},
    {{0, 0, 0},  {1, 0, 0},   {-1, 0, 0}, {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},  {0, 1, 1},  {0, -1, -1},
     {0, 1, -1}, {0, -1, 1},  {1, 0, 1},  {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 1, 0},  {-1, -1, 0}, {1, -1, 0},
     {-1, 1, 0}},
    {{0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},   {0, 0, -1},  {0, 1, 1},   {0, -1, -1},
     {0, 1, -1},  {0, -1, 1},  {1, 0, 1},   {-1, 0, -1}, {1, 0, -1},  {-1, 0, 1},  {1, 1, 0},   {-1, -1, 0}, {1, -1, 0},
     {-1, 1, 0},  {1, 1, 1},   {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1},  {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1}}};
static_assert(lbm::velocity_table_matches<lbm::D2Q9>(kTable[0]), "D2Q9 table");
static_assert(lbm::velocity_table_matches<lbm::D3Q19>(kTable[1]), "D3Q19 table");
static_assert(lbm::velocity_table_matches<lbm::D3Q27>(kTable[2]), "D3Q27 table");
}  // namespace
