/* host/coo2csc.h — COO -> compressed conversion, same signature as the reference's (final/coo2csc.h:5-13). */
#ifndef BSPGEMM_COO2CSC_H
#define BSPGEMM_COO2CSC_H
#include "../../include/bspgemm_host.h"
#endif
