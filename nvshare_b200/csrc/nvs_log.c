/* nvs_log.c -- the one definition of the debug switch (set from NVSHARE_DEBUG)
 * and of the "host process is exiting" flag, and the line writer. */
#include "nvs_log.h"

#include <stdarg.h>

void nvs_log_line(const char *level, const char *fmt, ...)
{
	char buf[1024];
	int n = snprintf(buf, sizeof(buf), "[NVSHARE][%s]: ", level);
	va_list ap;
	va_start(ap, fmt);
	int m = vsnprintf(buf + n, sizeof(buf) - (size_t)n - 1, fmt, ap);
	va_end(ap);
	if (m < 0)
		m = 0;
	n += m < (int)sizeof(buf) - n - 1 ? m : (int)sizeof(buf) - n - 2; /* (a longer line is cut, never overrun) */
	buf[n++] = '\n';
	fwrite(buf, 1, (size_t)n, stderr);
}

int nvs_debug_enabled = 0;
volatile int nvs_process_exiting = 0;
