/*
 * nvshare_wire.h -- the nvshare Unix-socket protocol, byte-compatible with the
 * reference so that our daemon, ctl and client library interoperate with the
 * reference's binaries in any combination.
 *
 * Reference interface this header replaces:
 *   src/comm.h:59-68   enum message_type (packed, 1 byte), 8 types
 *   src/comm.h:70-80   struct message (packed): type @0, pod_name[254] @1,
 *                      pod_namespace[254] @255, id (LE u64) @509, data[20] @517;
 *                      sizeof == 537.  Fixed-size frames, no length prefix.
 *   src/comm.h:45      socket directory "/var/run/nvshare/"
 *   src/comm.c:73-87   socket path = dir + "scheduler.sock"
 *   src/common.h:88    NVSHARE_UNREGISTERED_ID
 *   src/scheduler.c:591 / :338   id field of daemon messages (7331) / DROP_LOCK (1337)
 *   src/cli.c:80,103   id field of ctl messages (0xBEEF)
 *
 * Extension (wire-compatible; reference peers never look at `data` for these
 * types -- src/client.c:298-319, src/scheduler.c:464-494): our client may put
 * ASCII hints in `data` of REQ_LOCK / LOCK_RELEASED and our daemon may put
 * hints in `data` of DROP_LOCK / LOCK_OK.  See NVS_HINT_* below.
 */
#ifndef NVSHARE_WIRE_H
#define NVSHARE_WIRE_H

#include <stdint.h>
#include <stddef.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVS_MSG_DATA_LEN      20
#define NVS_POD_NAME_LEN      254
#define NVS_POD_NAMESPACE_LEN 254
#define NVS_MSG_SIZE          537

#define NVS_DEFAULT_SOCK_DIR  "/var/run/nvshare/"
#define NVS_SOCK_NAME         "scheduler.sock"
#define NVS_ENV_SOCK_DIR      "NVSHARE_SOCK_DIR" /* ours only: relocate the socket (tests, rootless) */
#define NVS_ENV_DEBUG         "NVSHARE_DEBUG"

#define NVS_UNREGISTERED_ID   0xF00DF00DF00DF00Dull
#define NVS_ID_DAEMON         7331ull
#define NVS_ID_DAEMON_TIMER   1337ull
#define NVS_ID_CTL            0xBEEFull
#define NVS_ID_CLIENT_PREREG  1234ull /* reference src/client.c:220 */

enum nvs_msg_type {
	NVS_REGISTER      = 1,
	NVS_SCHED_ON      = 2,
	NVS_SCHED_OFF     = 3,
	NVS_REQ_LOCK      = 4,
	NVS_LOCK_OK       = 5,
	NVS_DROP_LOCK     = 6,
	NVS_LOCK_RELEASED = 7,
	NVS_SET_TQ        = 8,
};

struct nvs_msg {
	uint8_t  type;
	char     pod_name[NVS_POD_NAME_LEN];
	char     pod_namespace[NVS_POD_NAMESPACE_LEN];
	uint64_t id;
	char     data[NVS_MSG_DATA_LEN];
} __attribute__((__packed__));

_Static_assert(sizeof(struct nvs_msg) == NVS_MSG_SIZE, "wire frame must be 537 bytes");
_Static_assert(offsetof(struct nvs_msg, pod_name) == 1, "pod_name offset");
_Static_assert(offsetof(struct nvs_msg, pod_namespace) == 255, "pod_namespace offset");
_Static_assert(offsetof(struct nvs_msg, id) == 509, "id offset");
_Static_assert(offsetof(struct nvs_msg, data) == 517, "data offset");

/*
 * `data` hints (ASCII, NUL-terminated, <= 19 chars).  Absent / unparsable
 * hints mean "no information" and select the reference-compatible behaviour.
 *   REQ_LOCK   "n<MiB>"  physical HBM the requester must map before it can run
 *   DROP_LOCK  "w<k>n<MiB>"  k clients wait behind the holder (k == 0: it may keep
 *                        its slabs resident); the next one must map <MiB>
 *   REQ_LOCK   "p<MiB>"  (from the client that is being granted / holds the lock)
 *                        it cannot map <MiB> more: memory pressure
 *   DROP_LOCK  "e<MiB>"  (to clients that do not hold the lock) evict at least <MiB>
 *   register reply: data[17] == '2' marks a daemon that speaks these hints; without
 *   it our client behaves like the conservative default (evict everything on release).
 */
#define NVS_HINT_NEED_PREFIX     'n'
#define NVS_HINT_WAITERS_PREFIX  'w'
#define NVS_HINT_PRESSURE_PREFIX 'p' /* REQ_LOCK "p<MiB>" from the client that is mapping: it is short of HBM */
#define NVS_HINT_EVICT_PREFIX    'e' /* DROP_LOCK "e<MiB>" to clients that do NOT hold the lock: release HBM */
#define NVS_CAP_MARKER           '2' /* register reply data[17]: the daemon understands the hints            */

/* helpers shared by the daemon, the CLI and the client library (not exported from libnvshare.so) */
#pragma GCC visibility push(hidden)
const char *nvs_msg_type_name(unsigned type);

/* Fills `out` (size >= 108) with the scheduler socket path; honours NVS_ENV_SOCK_DIR. */
int nvs_socket_path(char *out, size_t outlen);
/* Directory part of the above, with trailing '/'. */
int nvs_socket_dir(char *out, size_t outlen);
/* Path of the pinned-host pool shared by the clients of the scheduler listening on
 * the socket above: NVSHARE_POOL_PATH, else /dev/shm/nvshare-pool-<fnv1a(socket path)>. */
int nvs_pool_path(char *out, size_t outlen);

int nvs_listen(const char *path, int backlog);          /* -> nonblocking listening fd or -1 */
int nvs_accept(int lfd);                                /* -> nonblocking fd, -1 (errno set) */
int nvs_connect(const char *path);                      /* -> blocking fd or -1              */
ssize_t nvs_write_all(int fd, const void *buf, size_t n); /* n on success, -1 on error       */
ssize_t nvs_read_all(int fd, void *buf, size_t n);        /* bytes read (< n on EOF), -1     */
#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif
#endif /* NVSHARE_WIRE_H */
