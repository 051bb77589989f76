/*
 * ctl.c -- nvsharectl: one-shot control messages to nvshare-scheduler.
 *
 * Drop-in for the reference CLI (src/cli.c): same flags, same messages on the
 * wire, same user-visible text and exit codes (goldens recorded from the
 * reference binary in tests/golden/ctl_golden.json):
 *   -T n / --set-tq=n        SET_TQ, `data` = decimal n, id 0xBEEF   (cli.c:74-93)
 *   -S on|off / --anti-thrash=s   SCHED_ON / SCHED_OFF, id 0xBEEF    (cli.c:96-114)
 *   -h / --help, no action, or a stray positional: usage on stderr, exit 0
 *                                                                    (cli.c:175-183)
 *   anti-thrash is applied before the TQ when both are given         (cli.c:138-171)
 *
 * The reference links a vendored option parser (xopt); this file carries its
 * own 60-line argument scanner instead, which reproduces the parser's error
 * strings ("missing option value", "invalid option", ...), because scripts
 * match on them.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/nvshare_wire.h"
#include "nvs_log.h"

struct opt_def {
	char short_name;
	const char *long_name;
	const char *arg_name; /* NULL: flag */
	const char *help;
};

static const struct opt_def OPTS[] = {
	{'T', "set-tq", "n",
	 "Set the time quantum of the scheduler to TQ seconds. Only accepts positive integers."},
	{'S', "anti-thrash", "s",
	 "Set the desired status of the scheduler. Only accepts values \"on\" or \"off\"."},
	{'h', "help", NULL, "Shows this help message"},
};
#define N_OPTS (sizeof(OPTS) / sizeof(OPTS[0]))

struct settings {
	int tq;
	const char *anti_thrash;
	int help;
};

static void usage(FILE *f)
{
	fprintf(f, "usage: nvsharectl [options]\n\n"
		   "A command line utility to configure the nvshare scheduler.\n\n");
	for (size_t i = 0; i < N_OPTS; ++i) {
		char left[64];
		if (OPTS[i].arg_name)
			snprintf(left, sizeof(left), "-%c, --%s=%s", OPTS[i].short_name,
				 OPTS[i].long_name, OPTS[i].arg_name);
		else
			snprintf(left, sizeof(left), "-%c, --%s", OPTS[i].short_name, OPTS[i].long_name);
		/* option column padded so that help text starts 10 columns past the widest entry */
		fprintf(f, "%-29s%s\n", left, OPTS[i].help);
	}
	fprintf(f, "\n\n");
}

static void store(struct settings *cfg, const struct opt_def *o, const char *value, const char *shown_long,
		  int is_long)
{
	if (o->short_name == 'h') {
		cfg->help = 1;
	} else if (o->short_name == 'S') {
		cfg->anti_thrash = value;
	} else {
		char *end = NULL;
		long v = strtol(value, &end, 0);
		if (end == value || *end != '\0') {
			if (is_long)
				nvs_fatal("Error: value isn't a valid number: --%s=%s", shown_long, value);
			nvs_fatal("Error: value isn't a valid number: -%c %s", o->short_name, value);
		}
		cfg->tq = (int)v;
	}
}

static void parse(int argc, char **argv, struct settings *cfg)
{
	int only_positionals = 0;
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		if (only_positionals || a[0] != '-' || a[1] == '\0')
			continue; /* positionals are collected and ignored */
		if (strcmp(a, "--") == 0) {
			only_positionals = 1;
			continue;
		}
		if (a[1] == '-') { /* --name or --name=value */
			const char *name = a + 2;
			const char *eq = strchr(name, '=');
			size_t nlen = eq ? (size_t)(eq - name) : strlen(name);
			const struct opt_def *o = NULL;
			for (size_t k = 0; k < N_OPTS; ++k)
				if (strlen(OPTS[k].long_name) == nlen && strncmp(OPTS[k].long_name, name, nlen) == 0)
					o = &OPTS[k];
			if (!o)
				nvs_fatal("Error: invalid option: %s", a);
			if (!o->arg_name) {
				if (eq)
					nvs_fatal("Error: option doesn't take a value: %s", a);
				store(cfg, o, NULL, o->long_name, 1);
			} else if (eq) {
				store(cfg, o, eq + 1, o->long_name, 1);
			} else {
				if (i + 1 >= argc || argv[i + 1][0] == '-')
					nvs_fatal("Error: missing option value: %s", a);
				store(cfg, o, argv[++i], o->long_name, 1);
			}
			continue;
		}
		/* -x, or a cluster of flags ending in at most one option that takes a value */
		for (const char *p = a + 1; *p; ++p) {
			const struct opt_def *o = NULL;
			for (size_t k = 0; k < N_OPTS; ++k)
				if (OPTS[k].short_name == *p)
					o = &OPTS[k];
			if (!o)
				nvs_fatal("Error: invalid option: -%c", *p);
			if (!o->arg_name) {
				store(cfg, o, NULL, o->long_name, 0);
				continue;
			}
			if (p[1] != '\0')
				nvs_fatal("Error: short option parameters must be separated, not condensed: %s", a);
			if (i + 1 >= argc || argv[i + 1][0] == '-')
				nvs_fatal("Error: missing option value: -%c", *p);
			store(cfg, o, argv[++i], o->long_name, 0);
		}
	}
}

static int send_one(const struct nvs_msg *m, const char *path)
{
	int fd = nvs_connect(path);
	if (fd < 0) {
		nvs_info("Failed to connect to UNIX socket at %s\n", path);
		nvs_fatal("nvshare_connect() failed");
	}
	int rc = nvs_write_all(fd, m, sizeof(*m)) == (ssize_t)sizeof(*m) ? 0 : -1;
	nvs_must(close(fd) == 0);
	return rc;
}

int main(int argc, char **argv)
{
	struct settings cfg = {0, NULL, 0};
	char path[108];
	int actions = 0;

	parse(argc, argv, &cfg);
	if (nvs_socket_path(path, sizeof(path)) != 0)
		nvs_fatal("Failed to obtain nvshare-scheduler socket path.");

	if (cfg.anti_thrash != NULL) {
		int on;
		if (strcmp(cfg.anti_thrash, "on") == 0)
			on = 1;
		else if (strcmp(cfg.anti_thrash, "off") == 0)
			on = 0;
		else
			nvs_fatal("Invalid option for --anti-thrash (-S). Must be one of 'on' or 'off'.");
		struct nvs_msg m;
		memset(&m, 0, sizeof(m));
		m.type = on ? NVS_SCHED_ON : NVS_SCHED_OFF;
		m.id = NVS_ID_CTL;
		if (send_one(&m, path) != 0)
			nvs_info("Failed to turn the nvshare-scheduler %s.", cfg.anti_thrash);
		else
			nvs_info("Successfully turned the nvshare-scheduler %s.", cfg.anti_thrash);
		actions++;
	}

	if (cfg.tq != 0) {
		if (cfg.tq <= 0)
			nvs_fatal("Invalid option for --set-tq. TQ value must be a positive integer.");
		struct nvs_msg m;
		memset(&m, 0, sizeof(m));
		m.type = NVS_SET_TQ;
		m.id = NVS_ID_CTL;
		snprintf(m.data, sizeof(m.data), "%lld", (long long)cfg.tq);
		if (send_one(&m, path) != 0)
			nvs_info("Failed to set nvshare-scheduler TQ to %d seconds.", cfg.tq);
		else
			nvs_info("Successfully set the nvshare-scheduler TQ to %d seconds.", cfg.tq);
		actions++;
	}

	if (cfg.help || actions == 0) {
		usage(stderr);
		exit(0);
	}
	return 0;
}
