/* Compatibility header: code written against the reference's
 * include/iterative/savgolFilter.h compiles unchanged against libsavgol_b200. */
#ifndef SAVGOL_FILTER_H
#define SAVGOL_FILTER_H
#include "savgol_b200.h"
#endif
