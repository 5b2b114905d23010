/* Compatibility header for the reference's include/iterative/savgol2d.h. */
#ifndef SAVGOL2D_H
#define SAVGOL2D_H
#include "savgol_b200.h"
#endif
