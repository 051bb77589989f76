/* nvs_log.c -- the one definition of the debug switch (set from NVSHARE_DEBUG)
 * and of the "host process is exiting" flag. */
#include "nvs_log.h"

int nvs_debug_enabled = 0;
volatile int nvs_process_exiting = 0;
