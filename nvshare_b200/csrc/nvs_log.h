/*
 * nvs_log.h -- stderr logging with the reference's line prefix
 * ("[NVSHARE][LEVEL]: ", reference src/common.h:17-44) because operators and
 * the reference's README (README.md:284-356) grep scheduler/client logs for it.
 * A fatal log terminates the process with status 1, like the reference's
 * log_fatal / true_or_exit (src/common.h:24-28,47-52).
 */
#ifndef NVS_LOG_H
#define NVS_LOG_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <unistd.h>

extern int nvs_debug_enabled __attribute__((visibility("hidden")));
/* set once the host process has entered exit(): library threads must not call exit() again */
extern volatile int nvs_process_exiting __attribute__((visibility("hidden")));
#define NVS_EXITING() __atomic_load_n(&nvs_process_exiting, __ATOMIC_RELAXED)

/* One line = one write: the prefix, the text and the newline are put together first, so that lines of
 * different threads (the hook runs inside multi-threaded applications) never interleave and a daemon that
 * logs every frame pays one system call per line on its unbuffered stderr, not three. */
void nvs_log_line(const char *level, const char *fmt, ...) __attribute__((format(printf, 2, 3), visibility("hidden")));
#define nvs_log_at(level, ...) nvs_log_line(level, __VA_ARGS__)

#define nvs_info(...)  nvs_log_at("INFO", __VA_ARGS__)
#define nvs_warn(...)  nvs_log_at("WARN", __VA_ARGS__)
#define nvs_debug(...)                                       \
	do {                                                 \
		if (nvs_debug_enabled)                       \
			nvs_log_at("DEBUG", __VA_ARGS__);    \
	} while (0)
#define nvs_fatal(...)                                       \
	do {                                                 \
		nvs_log_at("FATAL", __VA_ARGS__);            \
		if (NVS_EXITING())                           \
			_exit(1);                            \
		exit(1);                                     \
	} while (0)
#define nvs_fatal_errno(...)                                 \
	do {                                                 \
		int e_ = errno;                              \
		nvs_log_at("FATAL", __VA_ARGS__);            \
		fprintf(stderr, "errno = %s\n", strerror(e_)); \
		exit(1);                                     \
	} while (0)

/* invariant check: failure is fatal for the hosting process (reference semantics) */
#define nvs_must(cond)                                       \
	do {                                                 \
		if (!(cond))                                 \
			nvs_fatal("Condition failed: %s", #cond); \
	} while (0)

#endif /* NVS_LOG_H */
