/*
 * gpu_ledger.c -- see gpu_ledger.h.  Host-side bookkeeping only: a ~50 KiB file in /dev/shm,
 * a robust process-shared mutex, a table of claims.  No CUDA calls here.
 */
#define _GNU_SOURCE
#include "gpu_ledger.h"

#include <errno.h>
#include <fcntl.h>
#include <inttypes.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include "nvs_log.h"

#define GL_MAGIC 0x6e7673684750554cull /* "nvshGPUL" */
#define GL_VERSION 2u /* 2: claims carry their owner's start time */
#define GL_MAX_DEVS 32
#define GL_MAX_CLAIMS 2048

struct gl_dev {
	uint8_t uuid[16];
	uint64_t total_bytes;
	uint32_t in_use;
	uint32_t pad;
};

struct gl_claim {
	int32_t pid; /* 0 = free entry */
	uint16_t dev;
	uint16_t kind;
	uint64_t bytes;
	uint64_t start; /* when that pid started (clock ticks since boot, /proc/<pid>/stat field 22): a pid
	                 * that has been recycled by an unrelated long-lived process must not keep the claim */
};

struct gl_file {
	volatile uint64_t magic;
	uint32_t version;
	uint32_t n_claims_max; /* GL_MAX_CLAIMS of whoever created the file */
	uint64_t pid_ns;       /* liveness of an owner is kill(pid, 0): one pid namespace only */
	pthread_mutex_t mu;
	struct gl_dev devs[GL_MAX_DEVS];
	struct gl_claim claims[GL_MAX_CLAIMS];
};

static struct gl_file *g_gl;       /* NULL: no usable ledger */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static double g_last_reap_ms;      /* process-local: when this process last looked for dead owners */

static double gl_now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static uint64_t gl_pid_ns(void)
{
	struct stat st;
	return stat("/proc/self/ns/pid", &st) == 0 ? (uint64_t)st.st_ino : 0;
}

/* Field 22 of /proc/<pid>/stat; 0 = no such process (or no /proc: then pids alone decide). */
static uint64_t gl_start_time(int32_t pid)
{
	char path[64], buf[1024];
	snprintf(path, sizeof(path), "/proc/%d/stat", (int)pid);
	int fd = open(path, O_RDONLY | O_CLOEXEC);
	if (fd < 0)
		return 0;
	ssize_t n = read(fd, buf, sizeof(buf) - 1);
	close(fd);
	if (n <= 0)
		return 0;
	buf[n] = '\0';
	char *p = strrchr(buf, ')'); /* the command name may contain anything, parentheses included */
	if (!p)
		return 0;
	p++;
	for (int field = 3; field < 22; ++field) { /* p is in front of field 3 (state) */
		while (*p == ' ')
			p++;
		while (*p && *p != ' ')
			p++;
	}
	return strtoull(p, NULL, 10);
}

static uint64_t g_my_start; /* of this process, re-read after a fork (see gl_me) */
static int32_t g_my_pid;

static int32_t gl_me(uint64_t *start)
{
	const int32_t me = (int32_t)getpid();
	if (me != g_my_pid) { /* first call, or a child of the process that opened the ledger */
		g_my_start = gl_start_time(me);
		g_my_pid = me;
	}
	if (start)
		*start = g_my_start;
	return me;
}

static void gl_lock(void)
{
	int r = pthread_mutex_lock(&g_gl->mu);
	if (r == EOWNERDEAD) /* its holder died inside: entries are independent words, nothing to repair */
		pthread_mutex_consistent(&g_gl->mu);
}

static void gl_unlock(void)
{
	pthread_mutex_unlock(&g_gl->mu);
}

/* drop the claims of owners that no longer exist (mutex held) */
static unsigned gl_reap_locked(void)
{
	unsigned n = 0;
	int32_t last = 0;
	int last_dead = 0;
	uint64_t last_start = 0;
	for (unsigned i = 0; i < GL_MAX_CLAIMS; ++i) {
		struct gl_claim *c = &g_gl->claims[i];
		if (c->pid <= 0)
			continue;
		if (c->pid != last) {
			last = c->pid;
			last_dead = kill(c->pid, 0) != 0 && errno == ESRCH;
			last_start = last_dead ? 0 : gl_start_time(c->pid);
		}
		/* gone, or the pid now belongs to somebody who started at another time */
		if (last_dead || (last_start && c->start && last_start != c->start)) {
			memset(c, 0, sizeof(*c));
			n++;
		}
	}
	g_last_reap_ms = gl_now_ms();
	return n;
}

static void gl_open_once(void)
{
	const char *env = getenv("NVSHARE_GPU_LEDGER");
	char path[256];
	if (env && (!strcmp(env, "off") || !strcmp(env, "0")))
		return;
	if (env && *env)
		snprintf(path, sizeof(path), "%s", env);
	else
		/* the layout version is part of the name: clients of two versions on one node keep two ledgers
		 * instead of the newer one finding a file it cannot use */
		snprintf(path, sizeof(path), "/dev/shm/nvshare-gpus-%u-v%u", (unsigned)geteuid(), GL_VERSION);

	int creator = 1;
	int fd = open(path, O_RDWR | O_CREAT | O_EXCL | O_CLOEXEC | O_NOFOLLOW, 0600);
	if (fd < 0 && errno == EEXIST) {
		creator = 0;
		fd = open(path, O_RDWR | O_CLOEXEC | O_NOFOLLOW);
	}
	if (fd < 0) {
		nvs_debug("gpu ledger: cannot open %s (%s): per-GPU accounting across processes is off", path, strerror(errno));
		return;
	}
	/* same rule as for the shared host pool: only a regular 0600 file of this very user */
	struct stat st;
	if (creator)
		(void)fchmod(fd, 0600);
	if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_uid != geteuid() || (st.st_mode & 07777) != 0600 ||
	    st.st_nlink != 1) {
		nvs_warn("gpu ledger: %s is not a ledger of ours (owner %d, mode %o): per-GPU accounting across processes is off",
			 path, (int)st.st_uid, (unsigned)(st.st_mode & 07777));
		close(fd);
		return;
	}
	if (creator && ftruncate(fd, (off_t)sizeof(struct gl_file)) != 0) {
		close(fd);
		unlink(path);
		return;
	}
	if (!creator) { /* the creator sizes the file, then publishes the magic */
		/* (a short file that nobody has touched for seconds is not being created: it is another layout) */
		for (int i = 0; i < 2000 && (fstat(fd, &st) != 0 || ((size_t)st.st_size < sizeof(struct gl_file) &&
									 time(NULL) - st.st_mtime < 3)); ++i)
			usleep(1000);
		if ((size_t)st.st_size != sizeof(struct gl_file)) {
			nvs_warn("gpu ledger: %s has an unexpected size (another version?): per-GPU accounting across processes is off", path);
			close(fd);
			return;
		}
	}
	struct gl_file *f = mmap(NULL, sizeof(struct gl_file), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (f == MAP_FAILED) {
		if (creator)
			unlink(path);
		return;
	}
	if (creator) {
		pthread_mutexattr_t at;
		pthread_mutexattr_init(&at);
		pthread_mutexattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
		pthread_mutexattr_setrobust(&at, PTHREAD_MUTEX_ROBUST);
		pthread_mutex_init(&f->mu, &at);
		pthread_mutexattr_destroy(&at);
		f->version = GL_VERSION;
		f->n_claims_max = GL_MAX_CLAIMS;
		f->pid_ns = gl_pid_ns();
		__atomic_store_n(&f->magic, GL_MAGIC, __ATOMIC_RELEASE);
	} else {
		for (int i = 0; i < 2000 && __atomic_load_n(&f->magic, __ATOMIC_ACQUIRE) != GL_MAGIC; ++i)
			usleep(1000);
		if (f->magic != GL_MAGIC || f->version != GL_VERSION || f->n_claims_max != GL_MAX_CLAIMS) {
			nvs_warn("gpu ledger: %s is not usable by this client (version): per-GPU accounting across processes is off", path);
			munmap(f, sizeof(struct gl_file));
			return;
		}
		if (f->pid_ns != gl_pid_ns()) {
			nvs_warn("gpu ledger: %s belongs to another pid namespace: per-GPU accounting across processes is off", path);
			munmap(f, sizeof(struct gl_file));
			return;
		}
	}
	g_gl = f;
	/* a recycled pid: whatever the table says about "us" was written by somebody who is gone */
	gl_lock();
	const int32_t me = gl_me(NULL);
	for (unsigned i = 0; i < GL_MAX_CLAIMS; ++i)
		if (g_gl->claims[i].pid == me)
			memset(&g_gl->claims[i], 0, sizeof(g_gl->claims[i]));
	gl_reap_locked();
	gl_unlock();
	nvs_debug("gpu ledger: %s (%s)", path, creator ? "created" : "attached");
}

static int gl_ready(int dev)
{
	return g_gl && dev >= 0 && dev < GL_MAX_DEVS;
}

int nvs_gl_device(const uint8_t uuid[16], uint64_t total_bytes)
{
	pthread_once(&g_once, gl_open_once);
	if (!g_gl)
		return -1;
	int idx = -1;
	gl_lock();
	for (int i = 0; i < GL_MAX_DEVS && idx < 0; ++i)
		if (g_gl->devs[i].in_use && !memcmp(g_gl->devs[i].uuid, uuid, 16))
			idx = i;
	for (int i = 0; i < GL_MAX_DEVS && idx < 0; ++i)
		if (!g_gl->devs[i].in_use) {
			memcpy(g_gl->devs[i].uuid, uuid, 16);
			g_gl->devs[i].in_use = 1;
			idx = i;
		}
	if (idx >= 0 && total_bytes)
		g_gl->devs[idx].total_bytes = total_bytes;
	gl_unlock();
	return idx;
}

/* this process's claim of `kind` on `dev`, created on demand (mutex held) */
static struct gl_claim *gl_mine_locked(int dev, int kind, int create)
{
	uint64_t my_start = 0;
	const int32_t me = gl_me(&my_start);
	struct gl_claim *spare = NULL;
	for (unsigned i = 0; i < GL_MAX_CLAIMS; ++i) {
		struct gl_claim *c = &g_gl->claims[i];
		if (c->pid == me && c->start && my_start && c->start != my_start)
			memset(c, 0, sizeof(*c)); /* left behind by an earlier owner of this pid */
		if (c->pid == me && c->dev == dev && c->kind == kind)
			return c;
		if (c->pid == 0 && !spare)
			spare = c;
	}
	if (!create)
		return NULL;
	if (!spare && gl_reap_locked())
		return gl_mine_locked(dev, kind, create);
	if (spare) {
		spare->pid = me;
		spare->start = my_start;
		spare->dev = (uint16_t)dev;
		spare->kind = (uint16_t)kind;
		spare->bytes = 0;
	}
	return spare;
}

static void gl_sums_locked(int dev, uint64_t *lent, uint64_t *max_own)
{
	uint64_t l = 0, m = 0;
	for (unsigned i = 0; i < GL_MAX_CLAIMS; ++i) {
		const struct gl_claim *c = &g_gl->claims[i];
		if (c->pid <= 0 || c->dev != dev)
			continue;
		if (c->kind == NVS_GL_LENT)
			l += c->bytes;
		else if (c->kind == NVS_GL_OWN && c->bytes > m)
			m = c->bytes;
	}
	*lent = l;
	*max_own = m;
}

int nvs_gl_lend(int dev, uint64_t bytes, uint64_t reserve_bytes)
{
	if (!gl_ready(dev))
		return 0;
	int rc = -1;
	gl_lock();
	for (int pass = 0; pass < 2 && rc != 0; ++pass) {
		uint64_t lent, max_own;
		gl_sums_locked(dev, &lent, &max_own);
		const uint64_t total = g_gl->devs[dev].total_bytes;
		const uint64_t spoken_for = reserve_bytes + max_own + lent;
		if (total == 0 || (spoken_for <= total && bytes <= total - spoken_for)) {
			struct gl_claim *c = gl_mine_locked(dev, NVS_GL_LENT, 1);
			if (c)
				c->bytes += bytes;
			rc = 0; /* a full table costs the accounting of this arena, never the arena */
		} else if (pass == 0 && gl_reap_locked() == 0) {
			break; /* nobody died: the refusal stands */
		}
	}
	gl_unlock();
	return rc;
}

static void gl_sub(int dev, int kind, uint64_t bytes)
{
	gl_lock();
	struct gl_claim *c = gl_mine_locked(dev, kind, 0);
	if (c) {
		c->bytes = c->bytes > bytes ? c->bytes - bytes : 0;
		if (c->bytes == 0)
			memset(c, 0, sizeof(*c));
	}
	gl_unlock();
}

void nvs_gl_return(int dev, uint64_t bytes)
{
	if (gl_ready(dev) && bytes)
		gl_sub(dev, NVS_GL_LENT, bytes);
}

void nvs_gl_own(int dev, int64_t delta)
{
	if (!gl_ready(dev) || delta == 0)
		return;
	if (delta < 0) {
		gl_sub(dev, NVS_GL_OWN, (uint64_t)-delta);
		return;
	}
	gl_lock();
	struct gl_claim *c = gl_mine_locked(dev, NVS_GL_OWN, 1);
	if (c)
		c->bytes += (uint64_t)delta;
	gl_unlock();
}

static void gl_query(int dev, uint64_t *lent, uint64_t *max_own)
{
	*lent = *max_own = 0;
	if (!gl_ready(dev))
		return;
	gl_lock();
	gl_sums_locked(dev, lent, max_own);
	/* a sum that shrinks somebody's cap must not contain the dead: look for them now and then */
	if ((*lent || *max_own) && gl_now_ms() - g_last_reap_ms > 1000.0 && gl_reap_locked())
		gl_sums_locked(dev, lent, max_own);
	gl_unlock();
}

uint64_t nvs_gl_lent(int dev)
{
	uint64_t l, m;
	gl_query(dev, &l, &m);
	return l;
}

uint64_t nvs_gl_max_own(int dev)
{
	uint64_t l, m;
	gl_query(dev, &l, &m);
	return m;
}

uint64_t nvs_gl_mine(int dev, int kind)
{
	if (!gl_ready(dev))
		return 0;
	gl_lock();
	struct gl_claim *c = gl_mine_locked(dev, kind, 0);
	uint64_t v = c ? c->bytes : 0;
	gl_unlock();
	return v;
}
