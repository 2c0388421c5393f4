// binary-spgemm_b200/csrc/tu_sort_w32.cu — k_fused_sort / k_fused_sort_async for ELL width 32 (see launch_sort.inl).
#define SORT_W 32
#include "launch_sort.inl"
