/* host/utils.h — readCOO, tic/toc, time statistics; same names as the reference's final/utils.h:7-13. */
#ifndef BSPGEMM_UTILS_H
#define BSPGEMM_UTILS_H
#include "../../include/bspgemm_host.h"
#endif
