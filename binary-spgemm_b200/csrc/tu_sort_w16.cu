// binary-spgemm_b200/csrc/tu_sort_w16.cu — k_fused_sort / k_fused_sort_async for ELL width 16 (see launch_sort.inl).
#define SORT_W 16
#include "launch_sort.inl"
