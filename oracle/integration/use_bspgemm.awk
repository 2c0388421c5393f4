# oracle/integration/use_bspgemm.awk — the whole source patch that switches the reference's drivers to libbspgemm.so.
#
#   awk -f use_bspgemm.awk final/SpGEMM_mpi_omp.c          > patched/SpGEMM_mpi_omp.c
#   awk -f use_bspgemm.awk final/SpGEMM_mpi_omp_validity.c > patched/SpGEMM_mpi_omp_validity.c
#
# It inserts four lines in front of `int test_mpi(` (final/SpGEMM_mpi_omp.c:294, final/SpGEMM_mpi_omp_validity.c:308), i.e.
# AFTER the reference's own definitions of SpGEMM_bigslice / SpGEMM_omp / SpGEMM_mpi (:15-225): those keep their names and
# stay what they are (in the validity driver SpGEMM_bigslice remains the serial CPU check, :337), while every call of
# SpGEMM_mpi that follows (:322, validity :331) becomes a call of bspgemm_SpGEMM_mpi — an UNDEFINED symbol of the
# executable that the dynamic linker resolves from libbspgemm.so.  (A macro placed before the definitions would rename the
# definitions too and the CPU code would keep running under the new name.)
/^int test_mpi\(/ {
  print "#ifdef USE_BSPGEMM                      /* hot path on B200: libbspgemm.so (include/bspgemm.h) */"
  print "#include \"bspgemm.h\""
  print "#define SpGEMM_mpi bspgemm_SpGEMM_mpi   /* call sites below only */"
  print "#endif"
}
{ print }
