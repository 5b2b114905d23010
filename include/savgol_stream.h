/* Compatibility header for the reference's include/iterative/savgol_stream.h. */
#ifndef SAVGOL_STREAM_H
#define SAVGOL_STREAM_H
#include "savgol_b200.h"
#endif
