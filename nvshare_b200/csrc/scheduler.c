/*
 * scheduler.c -- nvshare-scheduler: hands out the per-GPU lock, first come first
 * served, for at most TQ seconds at a time.
 *
 * Drop-in for the reference daemon (src/scheduler.c).  Observable behaviour
 * that is kept, with the reference lines it comes from:
 *   - socket dir 0711 + socket 0722 under /var/run/nvshare/      :536-547,588-589
 *   - REGISTER -> reply SCHED_ON|SCHED_OFF, id 7331, 16 hex digits of the new
 *     client id in `data`; a second REGISTER on the same connection closes it
 *                                                                 :159-206,402-410
 *   - REQ_LOCK / LOCK_RELEASED from an unregistered connection close it
 *                                                                 :464-494
 *   - FCFS queue, the head is the holder; duplicate REQ_LOCK ignored :123-137
 *   - LOCK_RELEASED from a waiter cancels its request             :139-155
 *   - TQ seconds after a grant (or a SET_TQ) the holder is sent ONE DROP_LOCK
 *     (id 1337), even when nobody is waiting                      :329-390
 *   - SCHED_OFF broadcasts, empties the queue; REQ_LOCK ignored while off;
 *     SCHED_ON broadcasts                                         :413-447
 *   - SET_TQ: `data` parsed with strtoll(base 0), no positivity check  :449-462
 *   - any receive error, EOF, short frame or hang-up removes the client and,
 *     if it held the lock, the next waiter is granted             :228-287,641-663
 *   - unknown message types are logged and ignored                :496-499
 *
 * What is different by design (not a port): one thread.  The reference runs an
 * epoll thread plus a timer thread sharing a mutex and a condition variable;
 * here the quantum is a timerfd in the same epoll set, so the state machine
 * has no locks and no cross-thread races.  Clients sit in an intrusive
 * doubly-linked list and the request queue is threaded through the same nodes,
 * so "already queued?" and removal are O(1) instead of list scans.
 *
 * Wire-compatible additions (include/nvshare_wire.h): REQ_LOCK may carry an
 * "n<MiB>" hint, DROP_LOCK carries "w<waiters>".  Reference peers ignore both.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <inttypes.h>
#include <limits.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/epoll.h>
#include <sys/ioctl.h>
#include <linux/sockios.h>
#include <sys/random.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/timerfd.h>

#include "../../include/nvshare_wire.h"
#include "nvs_log.h"

#define DEFAULT_TQ_SECONDS 30 /* reference src/scheduler.c:36 */
#define MAX_EVENTS         32 /* reference src/common.h:87    */
#define LISTEN_BACKLOG     32 /* reference src/comm.c:164     */

struct client {
	int fd;
	uint64_t id;
	uint64_t need_mib; /* last "n<MiB>" hint, 0 if none */
	int queued;
	int dead;                       /* dropped during this epoll batch; freed after it */
	struct client *prev, *next;     /* all connections (or the graveyard once dead)   */
	struct client *q_prev, *q_next; /* request queue (FCFS)   */
	char pod_name[NVS_POD_NAME_LEN];
	char pod_namespace[NVS_POD_NAMESPACE_LEN];
};

static struct client *all_head;
static struct client *graveyard; /* see drop_client() */
static struct client *q_head, *q_tail;
static unsigned q_len;
static int lock_held;
static int sched_on = 1;
static int tq = DEFAULT_TQ_SECONDS;
static int drop_sent;
static int ep_fd = -1, timer_fd = -1, listen_fd = -1;

static void try_schedule(void);

static const char *id_str(const struct client *c, char buf[24])
{
	if (c->id == NVS_UNREGISTERED_ID)
		strcpy(buf, "<UNREGISTERED>");
	else
		snprintf(buf, 24, "%016" PRIx64, c->id);
	return buf;
}

/* ------------------------------------------------------------- queue ---- */

static void q_push(struct client *c)
{
	c->queued = 1;
	c->q_next = NULL;
	c->q_prev = q_tail;
	if (q_tail)
		q_tail->q_next = c;
	else
		q_head = c;
	q_tail = c;
	q_len++;
}

/* Remove c's request.  If c was the head it was the holder: the lock is free. */
static void q_remove(struct client *c)
{
	if (!c->queued)
		return;
	if (q_head == c)
		lock_held = 0;
	if (c->q_prev)
		c->q_prev->q_next = c->q_next;
	else
		q_head = c->q_next;
	if (c->q_next)
		c->q_next->q_prev = c->q_prev;
	else
		q_tail = c->q_prev;
	c->queued = 0;
	c->q_prev = c->q_next = NULL;
	q_len--;
}

static void q_clear(void)
{
	while (q_head) {
		struct client *c = q_head;
		q_head = c->q_next;
		c->queued = 0;
		c->q_prev = c->q_next = NULL;
	}
	q_tail = NULL;
	q_len = 0;
	lock_held = 0;
}

/* ----------------------------------------------------------- clients ---- */

static void drop_client(struct client *c)
{
	char ids[24];
	nvs_info("Removing client %s", id_str(c, ids));
	q_remove(c);
	if (c->prev)
		c->prev->next = c->next;
	else
		all_head = c->next;
	if (c->next)
		c->next->prev = c->prev;
	epoll_ctl(ep_fd, EPOLL_CTL_DEL, c->fd, NULL);
	close(c->fd);
	/*
	 * Later events of the current epoll batch may still point at this node, so
	 * it is parked (not freed, so its address cannot be reused by a connection
	 * accepted in the same batch) and reaped when the batch is done.
	 */
	c->dead = 1;
	c->prev = NULL;
	c->next = graveyard;
	graveyard = c;
}

/*
 * An advisory frame (the forwarded memory-pressure hint) for a client that may be busy for
 * seconds inside a fetch or an eviction and not reading its socket.  Unlike the frames of
 * the reference protocol it may simply be dropped: the client under pressure repeats it
 * every second while it is stalled.  So it is only sent when the peer has consumed
 * everything sent before (nothing of ours is still queued on its socket) -- which both
 * coalesces the hints and keeps a slow but live client from ever filling its buffer with
 * them and being mistaken for a dead one.  Returns 0 sent, 1 skipped, -1 peer dead.
 */
static int send_frame(struct client *c, const struct nvs_msg *m);
static int send_advisory(struct client *c, const struct nvs_msg *m)
{
	int queued = 0;
	if (ioctl(c->fd, SIOCOUTQ, &queued) == 0 && queued > 0)
		return 1;
	return send_frame(c, m);
}

/* Strict like the reference: any failure to push a whole frame means the peer is dead. */
static int send_frame(struct client *c, const struct nvs_msg *m)
{
	char ids[24];
	ssize_t w;
	do {
		w = send(c->fd, m, sizeof(*m), MSG_NOSIGNAL | MSG_DONTWAIT);
	} while (w < 0 && errno == EINTR);
	if (w != (ssize_t)sizeof(*m)) {
		nvs_info("Failed to send message to client %s", id_str(c, ids));
		return -1;
	}
	nvs_info("Sent %s to client %s", nvs_msg_type_name(m->type), id_str(c, ids));
	return 0;
}

static void arm_timer(void)
{
	struct itimerspec its;
	memset(&its, 0, sizeof(its));
	if (tq > 0)
		its.it_value.tv_sec = tq;
	else
		its.it_value.tv_nsec = 1; /* tq <= 0: the quantum is already over */
	nvs_must(timerfd_settime(timer_fd, 0, &its, NULL) == 0);
	drop_sent = 0;
}

/* Grant the lock to the head of the queue; dead heads are discarded. */
static void try_schedule(void)
{
	struct nvs_msg m;
	memset(&m, 0, sizeof(m));
	m.type = NVS_LOCK_OK;
	m.id = NVS_ID_DAEMON;
	while (q_head) {
		if (send_frame(q_head, &m) != 0) {
			drop_client(q_head);
			continue;
		}
		lock_held = 1;
		arm_timer();
		return;
	}
	nvs_debug("try_schedule() called with no pending requests");
}

static void broadcast_status(void)
{
	struct nvs_msg m;
	memset(&m, 0, sizeof(m));
	m.type = sched_on ? NVS_SCHED_ON : NVS_SCHED_OFF;
	m.id = NVS_ID_DAEMON;
	for (struct client *c = all_head, *nx; c; c = nx) {
		nx = c->next;
		if (c->id == NVS_UNREGISTERED_ID)
			continue;
		if (send_frame(c, &m) != 0)
			drop_client(c);
	}
}

static uint64_t fresh_client_id(void)
{
	for (;;) {
		uint64_t id = 0;
		if (getrandom(&id, sizeof(id), 0) != (ssize_t)sizeof(id))
			id = ((uint64_t)rand() << 32) ^ (uint64_t)rand() ^ ((uint64_t)time(NULL) << 17);
		if (id == NVS_UNREGISTERED_ID)
			continue;
		int clash = 0;
		for (struct client *c = all_head; c; c = c->next)
			clash |= (c->id == id);
		if (!clash)
			return id;
	}
}

static void copy_field(char *dst, const char *src, size_t n)
{
	size_t i = 0;
	for (; i + 1 < n && src[i]; ++i)
		dst[i] = src[i];
	dst[i] = '\0';
}

/* ---------------------------------------------------------- messages ---- */

static void on_timer(void)
{
	uint64_t ticks;
	if (read(timer_fd, &ticks, sizeof(ticks)) < 0)
		return;
	nvs_debug("TQ elapsed");
	if (!lock_held || drop_sent || !q_head)
		return;
	struct nvs_msg m;
	memset(&m, 0, sizeof(m));
	m.type = NVS_DROP_LOCK;
	m.id = NVS_ID_DAEMON_TIMER;
	/* hint for our own clients (reference clients never read `data` here):
	 * how many clients wait behind the holder and how much HBM the next one must map */
	snprintf(m.data, sizeof(m.data), "%c%u%c%" PRIu64, NVS_HINT_WAITERS_PREFIX, q_len - 1, NVS_HINT_NEED_PREFIX,
		 q_head->q_next ? q_head->q_next->need_mib : 0);
	if (send_frame(q_head, &m) != 0) {
		drop_client(q_head);
		try_schedule();
	} else {
		drop_sent = 1;
	}
}

static void on_message(struct client *c, const struct nvs_msg *in)
{
	char ids[24];
	id_str(c, ids);

	switch (in->type) {
	case NVS_REGISTER: {
		nvs_info("Received %s", nvs_msg_type_name(in->type));
		if (c->id != NVS_UNREGISTERED_ID) {
			nvs_warn("Client %s is already registered", ids);
			drop_client(c);
			return;
		}
		c->id = fresh_client_id();
		copy_field(c->pod_name, in->pod_name, sizeof(c->pod_name));
		copy_field(c->pod_namespace, in->pod_namespace, sizeof(c->pod_namespace));
		struct nvs_msg m;
		memset(&m, 0, sizeof(m));
		m.type = sched_on ? NVS_SCHED_ON : NVS_SCHED_OFF;
		m.id = NVS_ID_DAEMON;
		snprintf(m.data, sizeof(m.data), "%016" PRIx64, c->id);
		m.data[17] = NVS_CAP_MARKER; /* after the NUL that ends the id: invisible to reference clients */
		if (send_frame(c, &m) != 0) {
			drop_client(c);
			return;
		}
		nvs_info("Registered client %016" PRIx64 " with Pod name = %s, Pod namespace = %s",
			 c->id, c->pod_name, c->pod_namespace);
		return;
	}
	case NVS_SCHED_ON:
		nvs_info("Received %s from %s", nvs_msg_type_name(in->type), ids);
		if (!sched_on) {
			sched_on = 1;
			nvs_info("Scheduler turned ON, broadcasting it...");
			broadcast_status();
		}
		return;
	case NVS_SCHED_OFF:
		nvs_info("Received %s from %s", nvs_msg_type_name(in->type), ids);
		if (sched_on) {
			nvs_info("Scheduler turned OFF, broadcasting it...");
			sched_on = 0;
			broadcast_status();
			/* every client now believes it owns the GPU: the queue is void */
			q_clear();
		}
		return;
	case NVS_SET_TQ: {
		nvs_info("Received %s from %s", nvs_msg_type_name(in->type), ids);
		char text[NVS_MSG_DATA_LEN + 1];
		memcpy(text, in->data, NVS_MSG_DATA_LEN);
		text[NVS_MSG_DATA_LEN] = '\0';
		char *end = NULL;
		errno = 0;
		long long v = strtoll(text, &end, 0);
		if (end != text && *end == '\0' && errno == 0) {
			tq = (int)v;
			arm_timer(); /* a new quantum starts now */
			nvs_info("New TQ = %d", tq);
		} else {
			nvs_info("Failed to parse new TQ from message");
		}
		return;
	}
	case NVS_REQ_LOCK:
		nvs_info("Received %s from %s", nvs_msg_type_name(in->type), ids);
		if (c->id == NVS_UNREGISTERED_ID) {
			drop_client(c);
			return;
		}
		if (!sched_on)
			return;
		if (in->data[0] == NVS_HINT_PRESSURE_PREFIX) {
			/* the sender cannot map its working set: ask everybody else to get out of HBM.
			 * Reference clients ignore a DROP_LOCK they do not hold the lock for.
			 * Only the client that has (been granted) the lock is in a position to press. */
			if (c != q_head) {
				nvs_debug("pressure hint from %s, which does not hold the lock: ignored", ids);
				return;
			}
			struct nvs_msg m;
			memset(&m, 0, sizeof(m));
			m.type = NVS_DROP_LOCK;
			m.id = NVS_ID_DAEMON_TIMER;
			snprintf(m.data, sizeof(m.data), "%c%" PRIu64, NVS_HINT_EVICT_PREFIX,
				 (uint64_t)strtoull(in->data + 1, NULL, 10));
			for (struct client *o = all_head, *nx; o; o = nx) {
				nx = o->next;
				if (o == c || o->id == NVS_UNREGISTERED_ID)
					continue;
				if (send_advisory(o, &m) < 0)
					drop_client(o);
			}
			return;
		}
		if (in->data[0] == NVS_HINT_NEED_PREFIX)
			c->need_mib = strtoull(in->data + 1, NULL, 10);
		if (c->queued) {
			if (in->data[0] != NVS_HINT_NEED_PREFIX) /* a refreshed need hint is not a mistake */
				nvs_warn("Client %s has already requested the lock", ids);
		} else {
			q_push(c);
		}
		if (!lock_held)
			try_schedule();
		return;
	case NVS_LOCK_RELEASED:
		nvs_info("Received %s from %s", nvs_msg_type_name(in->type), ids);
		if (c->id == NVS_UNREGISTERED_ID) {
			drop_client(c);
			return;
		}
		if (!sched_on)
			return;
		q_remove(c);
		if (!lock_held)
			try_schedule();
		return;
	default:
		nvs_info("Received message of unknown type %d from %s", (int)in->type, ids);
		return;
	}
}

static void on_client_event(struct client *c, uint32_t events)
{
	if (events & EPOLLIN) {
		struct nvs_msg in;
		ssize_t r;
		memset(&in, 0, sizeof(in));
		do {
			r = recv(c->fd, &in, sizeof(in), MSG_DONTWAIT);
		} while (r < 0 && errno == EINTR);
		if (r == (ssize_t)sizeof(in)) {
			on_message(c, &in);
			/* a handler may have dropped the holder (e.g. a protocol violation) */
			if (!lock_held && sched_on && q_head)
				try_schedule();
			return;
		}
		if (r == 0) {
			char ids[24];
			nvs_debug("Client %s has closed the connection", id_str(c, ids));
		}
		/* EOF, short frame or error: the peer is gone as far as we care */
	} else if (!(events & (EPOLLERR | EPOLLHUP))) {
		return;
	}
	drop_client(c);
	if (!lock_held && sched_on)
		try_schedule();
}

static void on_accept(void)
{
	for (;;) {
		int fd = nvs_accept(listen_fd);
		if (fd < 0) {
			if (errno == EAGAIN || errno == EWOULDBLOCK || errno == ECONNABORTED)
				return;
			nvs_fatal_errno("accept() failed non-transiently");
		}
		struct client *c = calloc(1, sizeof(*c));
		nvs_must(c != NULL);
		c->fd = fd;
		c->id = NVS_UNREGISTERED_ID;
		struct epoll_event ev = {.events = EPOLLIN, .data.ptr = c};
		if (epoll_ctl(ep_fd, EPOLL_CTL_ADD, fd, &ev) != 0) {
			nvs_warn("Couldn't add %d to the epoll interest list", fd);
			close(fd);
			free(c);
			continue;
		}
		c->next = all_head;
		if (all_head)
			all_head->prev = c;
		all_head = c;
	}
}

static char pool_path[256];

static void on_term(int sig)
{
	(void)sig;
	if (pool_path[0])
		unlink(pool_path); /* the clients' shared pinned-host pool dies with its scheduler */
	_exit(0);
}

int main(void)
{
	char dir[108], path[108];

	signal(SIGPIPE, SIG_IGN);
	signal(SIGTERM, on_term);
	signal(SIGINT, on_term);
	if (getenv(NVS_ENV_DEBUG)) {
		nvs_debug_enabled = 1;
		nvs_info("nvshare-scheduler started in debug mode");
	} else {
		nvs_info("nvshare-scheduler started in normal mode");
	}

	if (nvs_socket_dir(dir, sizeof(dir)) != 0 || nvs_socket_path(path, sizeof(path)) != 0)
		nvs_fatal("socket path too long");
	if (mkdir(dir, 0711) != 0 && errno != EEXIST)
		nvs_fatal("Could not create scheduler socket directory %s", dir);
	if (chmod(dir, 0711) != 0)
		nvs_fatal("chmod() failed for %s", dir);

	nvs_must((ep_fd = epoll_create1(EPOLL_CLOEXEC)) >= 0);
	nvs_must((timer_fd = timerfd_create(CLOCK_MONOTONIC, TFD_NONBLOCK | TFD_CLOEXEC)) >= 0);
	if ((listen_fd = nvs_listen(path, LISTEN_BACKLOG)) < 0)
		nvs_fatal_errno("Failed to bind UNIX socket to %s", path);
	if (chmod(path, 0722) != 0)
		nvs_fatal("chmod() failed for %s", path);

	/* tags in data.ptr: NULL = listener, &timer_fd = quantum timer, else client */
	struct epoll_event ev = {.events = EPOLLIN, .data.ptr = NULL};
	nvs_must(epoll_ctl(ep_fd, EPOLL_CTL_ADD, listen_fd, &ev) == 0);
	ev.data.ptr = &timer_fd;
	nvs_must(epoll_ctl(ep_fd, EPOLL_CTL_ADD, timer_fd, &ev) == 0);

	/* a pool file left by a previous instance belongs to clients that are gone */
	if (nvs_pool_path(pool_path, sizeof(pool_path)) == 0)
		unlink(pool_path);
	else
		pool_path[0] = '\0';

	nvs_info("nvshare-scheduler listening on %s", path);

	for (;;) {
		struct epoll_event evs[MAX_EVENTS];
		int n = epoll_wait(ep_fd, evs, MAX_EVENTS, -1);
		if (n < 0) {
			if (errno == EINTR)
				continue;
			nvs_fatal_errno("epoll_wait() failed");
		}
		for (int i = 0; i < n; ++i) {
			void *tag = evs[i].data.ptr;
			if (tag == NULL)
				on_accept();
			else if (tag == &timer_fd)
				on_timer();
			else if (!((struct client *)tag)->dead)
				on_client_event(tag, evs[i].events);
		}
		while (graveyard) {
			struct client *c = graveyard;
			graveyard = c->next;
			free(c);
		}
	}
	return 1;
}
