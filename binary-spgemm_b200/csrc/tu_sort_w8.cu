// binary-spgemm_b200/csrc/tu_sort_w8.cu — k_fused_sort / k_fused_sort_async for ELL width 8 (see launch_sort.inl).
#define SORT_W 8
#include "launch_sort.inl"
