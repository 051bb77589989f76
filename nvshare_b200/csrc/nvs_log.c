/* nvs_log.c -- the one definition of the debug switch (set from NVSHARE_DEBUG). */
#include "nvs_log.h"

int nvs_debug_enabled = 0;
