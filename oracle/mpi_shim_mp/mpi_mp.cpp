// mpi_mp.cpp -- the 12 MPI calls of the reference over shared-memory mailboxes (see mpi.h).  TEST INFRASTRUCTURE ONLY.
//
// Environment (set by oracle/mpirun_shim.py): MPISHIM_RANK, MPISHIM_SIZE, MPISHIM_TAG (name stem of the /dev/shm objects).
// Control segment /dev/shm/<tag>.ctl: K x K mailboxes {state, bytes}.  Payload of mailbox (s -> d): file /dev/shm/<tag>.<s>.<d>.
#include "mpi.h"

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace {
struct Box { std::atomic<int> full; long long bytes; char pad[48]; };
int g_rank = 0, g_size = 1;
std::string g_tag;
Box *g_box = nullptr;           // [size][size]: g_box[s * size + d]

size_t width(MPI_Datatype t) { return t == MPI_BYTE ? 1u : 4u; }
[[noreturn]] void die(const char *what) { fprintf(stderr, "mpi shim (rank %d): %s\n", g_rank, what); _exit(3); }
std::string box_path(int s, int d) { return "/dev/shm/" + g_tag + "." + std::to_string(s) + "." + std::to_string(d); }

void wait_for(std::atomic<int> &f, int want)
{
	for (unsigned spin = 0; f.load(std::memory_order_acquire) != want; spin++) {
		if (spin > 2000) { sched_yield(); if (spin > 20000) usleep(50); }
	}
}

void send_to(int d, const void *buf, size_t bytes)
{
	Box &b = g_box[g_rank * g_size + d];
	wait_for(b.full, 0);
	if (bytes) {
		int fd = open(box_path(g_rank, d).c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0600);
		if (fd < 0) die("cannot open a mailbox for writing");
		const char *p = static_cast<const char *>(buf);
		size_t left = bytes;
		while (left) { ssize_t w = write(fd, p, left); if (w <= 0) die("mailbox write failed"); p += w; left -= (size_t) w; }
		close(fd);
	}
	b.bytes = (long long) bytes;
	b.full.store(1, std::memory_order_release);
}

size_t recv_from(int s, void *buf, size_t cap)
{
	Box &b = g_box[s * g_size + g_rank];
	wait_for(b.full, 1);
	const size_t bytes = (size_t) b.bytes;
	if (bytes > cap) die("message longer than the receive buffer");
	if (bytes) {
		int fd = open(box_path(s, g_rank).c_str(), O_RDONLY);
		if (fd < 0) die("cannot open a mailbox for reading");
		char *p = static_cast<char *>(buf);
		size_t left = bytes;
		while (left) { ssize_t r = read(fd, p, left); if (r <= 0) die("mailbox read failed"); p += r; left -= (size_t) r; }
		close(fd);
	}
	b.full.store(0, std::memory_order_release);
	return bytes;
}
}  // namespace

int MPI_Init(int *, char ***)
{
	const char *r = getenv("MPISHIM_RANK"), *s = getenv("MPISHIM_SIZE"), *t = getenv("MPISHIM_TAG");
	if (!r || !s || !t) { g_rank = 0; g_size = 1; return 0; }          // started by hand: one rank
	g_rank = atoi(r); g_size = atoi(s); g_tag = t;
	if (g_size < 1 || g_rank < 0 || g_rank >= g_size) die("bad MPISHIM_RANK / MPISHIM_SIZE");
	if (g_size == 1) return 0;
	int fd = open(("/dev/shm/" + g_tag + ".ctl").c_str(), O_RDWR);
	if (fd < 0) die("control segment missing (start the ranks with oracle/mpirun_shim.py)");
	const size_t bytes = sizeof(Box) * (size_t) g_size * g_size;
	void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (m == MAP_FAILED) die("cannot map the control segment");
	g_box = static_cast<Box *>(m);
	return 0;
}

int MPI_Finalize(void) { if (g_size > 1) MPI_Barrier(MPI_COMM_WORLD); return 0; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = g_size; return 0; }

int MPI_Barrier(MPI_Comm)
{
	if (g_size == 1) return 0;
	char c = 0;
	if (g_rank == 0) {
		for (int p = 1; p < g_size; p++) recv_from(p, &c, 1);
		for (int p = 1; p < g_size; p++) send_to(p, &c, 1);
	} else {
		send_to(0, &c, 1);
		recv_from(0, &c, 1);
	}
	return 0;
}

int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm)
{
	if (g_size == 1 || n <= 0) return 0;
	const size_t bytes = (size_t) n * width(t);
	if (g_rank == root) { for (int p = 0; p < g_size; p++) if (p != root) send_to(p, buf, bytes); }
	else recv_from(root, buf, bytes);
	return 0;
}

int MPI_Gatherv(const void *snd, int n, MPI_Datatype t, void *rcv, const int *counts, const int *displ, MPI_Datatype rt, int root, MPI_Comm)
{
	const size_t w = width(t), rw = width(rt);
	if (g_rank == root) {
		for (int p = 0; p < g_size; p++) {
			char *dst = static_cast<char *>(rcv) + (size_t) displ[p] * rw;
			if (p == root) { if (n > 0 && dst != snd) memcpy(dst, snd, (size_t) n * w); }
			else recv_from(p, dst, (size_t) counts[p] * rw);
		}
	} else send_to(root, snd, n > 0 ? (size_t) n * w : 0);
	return 0;
}

int MPI_Gather(const void *snd, int n, MPI_Datatype t, void *rcv, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{
	std::vector<int> counts(g_size, rn), displ(g_size);
	for (int p = 0; p < g_size; p++) displ[p] = p * rn;
	return MPI_Gatherv(snd, n, t, rcv, counts.data(), displ.data(), rt, root, c);
}

int MPI_Reduce(const void *snd, void *rcv, int n, MPI_Datatype t, MPI_Op, int root, MPI_Comm)
{
	// the reference only sums ints (src/MergerTree.cpp:645-646, src/Communication.cpp:325-326)
	if (t != MPI_INT) die("MPI_Reduce: only MPI_INT / MPI_SUM is implemented");
	if (g_size == 1) { if (n > 0 && rcv != snd) memcpy(rcv, snd, (size_t) n * 4); return 0; }
	if (g_rank == root) {
		std::vector<int> acc(static_cast<const int *>(snd), static_cast<const int *>(snd) + n), tmp(n);
		for (int p = 0; p < g_size; p++) {
			if (p == root) continue;
			recv_from(p, tmp.data(), (size_t) n * 4);
			for (int i = 0; i < n; i++) acc[i] += tmp[i];
		}
		memcpy(rcv, acc.data(), (size_t) n * 4);
	} else send_to(root, snd, (size_t) n * 4);
	return 0;
}

int MPI_Sendrecv(const void *snd, int n, MPI_Datatype t, int dest, int, void *rcv, int rn, MPI_Datatype rt, int source, int, MPI_Comm, MPI_Status *)
{
	if (dest == g_rank && source == g_rank) { if (n > 0 && rcv != snd) memcpy(rcv, snd, (size_t) n * width(t)); return 0; }
	send_to(dest, snd, n > 0 ? (size_t) n * width(t) : 0);
	recv_from(source, rcv, rn > 0 ? (size_t) rn * width(rt) : 0);
	return 0;
}

int MPI_Pack(const void *in, int n, MPI_Datatype t, void *out, int, int *pos, MPI_Comm)
{ const size_t b = (size_t) n * width(t); memcpy(static_cast<char *>(out) + *pos, in, b); *pos += (int) b; return 0; }

int MPI_Unpack(const void *in, int, int *pos, void *out, int n, MPI_Datatype t, MPI_Comm)
{ const size_t b = (size_t) n * width(t); memcpy(out, static_cast<const char *>(in) + *pos, b); *pos += (int) b; return 0; }
