/*
 * pingpong.c -- BASELINE config #1 (CPU only): K scripted clients hammer an
 * nvshare-scheduler with REQ_LOCK -> (LOCK_OK) -> LOCK_RELEASED cycles over the
 * 537-byte protocol; reports lock hand-offs per second and the REQ_LOCK -> LOCK_OK
 * latency distribution.  Works against the reference daemon and ours (same wire).
 *
 *   pingpong <socket path> <clients> <cycles per client>
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/socket.h>
#include <sys/un.h>

#include "../include/nvshare_wire.h"

static const char *sock_path;
static long cycles;
static double *lat_us; /* [client][cycle] */

static double now_us(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static void xfer(int fd, void *buf, int wr)
{
	size_t done = 0;
	while (done < NVS_MSG_SIZE) {
		ssize_t r = wr ? write(fd, (char *)buf + done, NVS_MSG_SIZE - done) : read(fd, (char *)buf + done, NVS_MSG_SIZE - done);
		if (r <= 0) {
			perror("socket");
			exit(1);
		}
		done += (size_t)r;
	}
}

static void *client(void *arg)
{
	long id = (long)arg;
	struct sockaddr_un a;
	memset(&a, 0, sizeof a);
	a.sun_family = AF_UNIX;
	strncpy(a.sun_path, sock_path, sizeof(a.sun_path) - 1);
	int fd = socket(AF_UNIX, SOCK_STREAM, 0);
	if (fd < 0 || connect(fd, (struct sockaddr *)&a, sizeof a) != 0) {
		perror("connect");
		exit(1);
	}
	struct nvs_msg m, in;
	memset(&m, 0, sizeof m);
	m.type = NVS_REGISTER;
	xfer(fd, &m, 1);
	xfer(fd, &in, 0);
	for (long c = 0; c < cycles; ++c) {
		m.type = NVS_REQ_LOCK;
		double t0 = now_us();
		xfer(fd, &m, 1);
		do {
			xfer(fd, &in, 0);
		} while (in.type != NVS_LOCK_OK);
		lat_us[id * cycles + c] = now_us() - t0;
		m.type = NVS_LOCK_RELEASED;
		xfer(fd, &m, 1);
	}
	close(fd);
	return NULL;
}

static int cmp(const void *a, const void *b)
{
	double x = *(const double *)a, y = *(const double *)b;
	return x < y ? -1 : x > y;
}

int main(int argc, char **argv)
{
	if (argc < 4)
		return 2;
	sock_path = argv[1];
	long k = atol(argv[2]);
	cycles = atol(argv[3]);
	lat_us = calloc((size_t)(k * cycles), sizeof(double));
	pthread_t th[64];
	double t0 = now_us();
	for (long i = 0; i < k; ++i)
		pthread_create(&th[i], NULL, client, (void *)i);
	for (long i = 0; i < k; ++i)
		pthread_join(th[i], NULL);
	double secs = (now_us() - t0) * 1e-6;
	long n = k * cycles;
	qsort(lat_us, (size_t)n, sizeof(double), cmp);
	printf("{\"clients\":%ld,\"handoffs\":%ld,\"seconds\":%.3f,\"handoffs_per_s\":%.0f,\"req_to_lock_ok_us\":{\"p50\":%.1f,"
	       "\"p99\":%.1f,\"max\":%.1f}}\n", k, n, secs, n / secs, lat_us[n / 2], lat_us[(long)(n * 0.99)], lat_us[n - 1]);
	return 0;
}
