// shift_table_inst.cu -- instantiates shift_table_kernel<S, SH, UMPA_INST_NW> for S = 3..19.
// Compiled once per window half-width (build.py passes -DUMPA_INST_NW=-1,0,...,6).
#include "shift_table.cuh"

#ifndef UMPA_INST_NW
#error "compile with -DUMPA_INST_NW=<-1..6>"
#endif

#define UMPA_CAT2(a, b) a##b
#define UMPA_CAT(a, b) UMPA_CAT2(a, b)
#if UMPA_INST_NW < 0
#define UMPA_ST_NAME shift_table_launch_plain
#else
#define UMPA_ST_NAME UMPA_CAT(shift_table_launch_nw, UMPA_INST_NW)
#endif

int UMPA_ST_NAME(UMPA_ST_ARGS) { return shift_table::dispatch_shift_table_s<UMPA_INST_NW>(S, a, b, p, grid, nt, smem, st); }
