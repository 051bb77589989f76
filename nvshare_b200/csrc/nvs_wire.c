/*
 * nvs_wire.c -- socket plumbing for the 537-byte nvshare protocol
 * (include/nvshare_wire.h).  Behavioural contract taken from the reference's
 * src/comm.c:73-227 and src/common.c:76-109: stream AF_UNIX socket, stale
 * socket file unlinked before bind, listen backlog 32, accepted fds are
 * non-blocking, EINTR is always retried, whole-frame helpers loop on short
 * transfers.  Unlike the reference, writes use MSG_NOSIGNAL so a dead peer
 * yields EPIPE instead of killing the process (SURVEY section 5 "latent").
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/socket.h>
#include <sys/un.h>

#include "../../include/nvshare_wire.h"
#include "nvs_log.h"

const char *nvs_msg_type_name(unsigned type)
{
	static const char *const names[] = {
		"UNKNOWN", "REGISTER", "SCHED_ON", "SCHED_OFF", "REQ_LOCK",
		"LOCK_OK", "DROP_LOCK", "LOCK_RELEASED", "SET_TQ",
	};
	return type < sizeof(names) / sizeof(names[0]) ? names[type] : names[0];
}

int nvs_socket_dir(char *out, size_t outlen)
{
	const char *dir = getenv(NVS_ENV_SOCK_DIR);
	if (dir == NULL || dir[0] == '\0')
		dir = NVS_DEFAULT_SOCK_DIR;
	size_t len = strlen(dir);
	int need_slash = (dir[len - 1] != '/');
	if (len + need_slash + 1 > outlen)
		return -1;
	memcpy(out, dir, len);
	if (need_slash)
		out[len++] = '/';
	out[len] = '\0';
	return 0;
}

int nvs_socket_path(char *out, size_t outlen)
{
	if (nvs_socket_dir(out, outlen) != 0)
		return -1;
	size_t len = strlen(out);
	if (len + sizeof(NVS_SOCK_NAME) > outlen)
		return -1;
	memcpy(out + len, NVS_SOCK_NAME, sizeof(NVS_SOCK_NAME));
	return 0;
}

int nvs_pool_path(char *out, size_t outlen)
{
	const char *forced = getenv("NVSHARE_POOL_PATH");
	if (forced && *forced)
		return snprintf(out, outlen, "%s", forced) < (int)outlen ? 0 : -1;
	char sock[108];
	if (nvs_socket_path(sock, sizeof(sock)) != 0)
		return -1;
	uint64_t h = 0xcbf29ce484222325ull; /* FNV-1a over the socket path: one pool per scheduler */
	for (const char *c = sock; *c; ++c)
		h = (h ^ (unsigned char)*c) * 0x100000001b3ull;
	return snprintf(out, outlen, "/dev/shm/nvshare-pool-%016llx", (unsigned long long)h) < (int)outlen ? 0 : -1;
}

static int fill_addr(struct sockaddr_un *addr, const char *path)
{
	memset(addr, 0, sizeof(*addr));
	addr->sun_family = AF_UNIX;
	if (strlen(path) >= sizeof(addr->sun_path)) {
		errno = ENAMETOOLONG;
		return -1;
	}
	strcpy(addr->sun_path, path);
	return 0;
}

int nvs_listen(const char *path, int backlog)
{
	struct sockaddr_un addr;
	if (fill_addr(&addr, path) != 0)
		return -1;
	int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_NONBLOCK | SOCK_CLOEXEC, 0);
	if (fd < 0)
		return -1;
	if (unlink(path) != 0 && errno != ENOENT)
		goto fail;
	if (bind(fd, (struct sockaddr *)&addr, sizeof(addr)) != 0)
		goto fail;
	if (listen(fd, backlog) != 0)
		goto fail;
	return fd;
fail: {
	int e = errno;
	close(fd);
	errno = e;
	return -1;
}
}

int nvs_accept(int lfd)
{
	int fd;
	do {
		fd = accept4(lfd, NULL, NULL, SOCK_NONBLOCK | SOCK_CLOEXEC);
	} while (fd < 0 && errno == EINTR);
	return fd;
}

int nvs_connect(const char *path)
{
	struct sockaddr_un addr;
	if (fill_addr(&addr, path) != 0)
		return -1;
	int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
	if (fd < 0)
		return -1;
	int rc;
	do {
		rc = connect(fd, (struct sockaddr *)&addr, sizeof(addr));
	} while (rc != 0 && errno == EINTR);
	if (rc != 0) {
		int e = errno;
		close(fd);
		errno = e;
		return -1;
	}
	return fd;
}

ssize_t nvs_write_all(int fd, const void *buf, size_t n)
{
	const char *p = buf;
	size_t left = n;
	while (left > 0) {
		ssize_t w = send(fd, p, left, MSG_NOSIGNAL);
		if (w < 0) {
			if (errno == EINTR)
				continue;
			return -1;
		}
		p += w;
		left -= (size_t)w;
	}
	return (ssize_t)n;
}

ssize_t nvs_read_all(int fd, void *buf, size_t n)
{
	char *p = buf;
	size_t got = 0;
	while (got < n) {
		ssize_t r = read(fd, p + got, n - got);
		if (r < 0) {
			if (errno == EINTR)
				continue;
			return -1;
		}
		if (r == 0)
			break;
		got += (size_t)r;
	}
	return (ssize_t)got;
}
