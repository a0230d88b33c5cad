// emd_tuning.cu -- the one place the EMD_* environment is read (emd_kernels.h, struct Tuning).
#include "emd_kernels.h"

#include <cstdlib>
#include <cstring>

namespace emd {
namespace {

struct Entry { const char* name; const char* env; int Tuning::*field; bool disable; };
// option name (emd_set_option) | environment variable | field | the variable DISABLES the feature when set to 1
const Entry kTable[] = {
    {"umma", "EMD_DISABLE_UMMA", &Tuning::umma, true},
    {"fused", "EMD_DISABLE_FUSED", &Tuning::fused, true},
    {"tma", "EMD_DISABLE_TMA", &Tuning::tma, true},
    {"pair", "EMD_DISABLE_PAIR", &Tuning::pair, true},
    {"final_umma", "EMD_DISABLE_FINAL_UMMA", &Tuning::final_umma, true},
    {"pdl", "EMD_DISABLE_PDL", &Tuning::pdl, true},
    {"graphs", "EMD_DISABLE_GRAPH", &Tuning::graphs, true},
    {"sliced_io", "EMD_DISABLE_SLICED_IO", &Tuning::sliced_io, true},
    {"halves", "EMD_DISABLE_HALVES", &Tuning::halves, true},
    {"mid_graph", "EMD_DISABLE_MID_GRAPH", &Tuning::mid_graph, true},
    {"skip_taps", "EMD_DISABLE_SKIP_TAPS", &Tuning::skip_taps, true},
    {"dw_tile", "EMD_DISABLE_DW_TILE", &Tuning::dw_tile, true},
    {"dw_reg", "EMD_DISABLE_DW_REG", &Tuning::dw_reg, true},
    {"dw_reg_all", "EMD_DW_REG_ALL", &Tuning::dw_reg_all, false},
    {"dw_strip", "EMD_DISABLE_DW_STRIP", &Tuning::dw_strip, true},
    {"dw_cols", "EMD_DISABLE_DW_COLS", &Tuning::dw_cols, true},
    {"pad_pitch", "EMD_DISABLE_PAD_PITCH", &Tuning::pad_pitch, true},
    {"fork_sms", "EMD_FORK_SMS", &Tuning::fork_sms, false},
    {"poison", "EMD_POISON", &Tuning::poison, false},
    {"strict", "EMD_STRICT", &Tuning::strict, false},
    {"graph_max_n", "EMD_GRAPH_MAX_N", &Tuning::graph_max_n, false},
    {"pair_min_rows", "EMD_PAIR_MIN_ROWS", &Tuning::pair_min_rows, false},
    {"pair_min_items", "EMD_PAIR_MIN_ITEMS", &Tuning::pair_min_items, false},
    {"io_slices", "EMD_IO_SLICES", &Tuning::io_slices, false},
    {"io_parts", "EMD_IO_PARTS", &Tuning::io_parts, false},
    {"dw_stages", "EMD_DW_STAGES", &Tuning::dw_stages, false},
    {"dw_sa", "EMD_DW_SA", &Tuning::dw_sa, false},
    {"dw_sb", "EMD_DW_SB", &Tuning::dw_sb, false},
    {"dw_sh", "EMD_DW_SH", &Tuning::dw_sh, false},
    {"dw_ring", "EMD_DW_RING", &Tuning::dw_ring, false},
};

Tuning from_env() {
  Tuning t;
  for (const Entry& e : kTable) {
    const char* v = getenv(e.env);
    if (!v || !*v) continue;
    if (e.disable) { if (v[0] != '0') t.*(e.field) = 0; }
    else t.*(e.field) = atoi(v);
  }
  return t;
}

}  // namespace

Tuning& tuning() {
  static Tuning t = from_env();
  return t;
}

bool tuning_set(const char* name, long long value) {
  for (const Entry& e : kTable)
    if (!strcmp(e.name, name)) { tuning().*(e.field) = (int)value; return true; }
  return false;
}

bool tuning_get(const char* name, long long* value) {
  for (const Entry& e : kTable)
    if (!strcmp(e.name, name)) { *value = tuning().*(e.field); return true; }
  return false;
}

int& last_launch_kind() {
  static thread_local int k = LK_NONE;
  return k;
}

}  // namespace emd
