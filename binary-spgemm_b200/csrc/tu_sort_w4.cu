// binary-spgemm_b200/csrc/tu_sort_w4.cu — k_fused_sort / k_fused_sort_async for ELL width 4 (see launch_sort.inl).
#define SORT_W 4
#include "launch_sort.inl"
