// main.cpp -- MetroCPP-b200 [--ranks K] <config.cfg>: the reference's driver loop (reference src/main.cpp:48-328)
// with the matching path on the GPU.  Snapshot 0 is the latest catalogue; the loop walks backwards
// in time: read the older catalogue, match the pair, clean, write the .mtree of the later snapshot,
// shift 1 -> 0.
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "IOSettings.h"
#include "MergerTree.h"
#include "global_vars.h"

static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- ranks ----------------------------------------------------------------------------------------
// The reference is `mpirun -n K MetroCPP cfg` (src/main.cpp:57-63: MPI_Comm_rank / MPI_Comm_size).  Here a rank is a process that
// drives one GPU; rank and world come from the environment any launcher sets (METRO_RANK/METRO_WORLD, torchrun's RANK/WORLD_SIZE/
// LOCAL_RANK, Open MPI's OMPI_COMM_WORLD_*), or from `--ranks K`, which makes this program its own launcher: it forks K copies of
// itself.  The ranks only ever meet inside the library (NCCL); the 128-byte NCCL id travels through a file in pathOutput.
static int EnvInt(const char *const *names, int dflt)
{
	for (; *names; names++) if (const char *v = getenv(*names)) return atoi(v);
	return dflt;
}

static int LaunchRanks(int ranks, int argc, char **argv)
{
	std::vector<pid_t> kids;
	const std::string tag = std::to_string((long long) getpid());
	for (int r = 0; r < ranks; r++) {
		pid_t pid = fork();
		if (pid < 0) { perror("fork"); return 2; }
		if (pid == 0) {
			setenv("METRO_RANK", std::to_string(r).c_str(), 1);
			setenv("METRO_WORLD", std::to_string(ranks).c_str(), 1);
			setenv("METRO_LOCAL_RANK", std::to_string(r).c_str(), 1);
			setenv("METRO_RUN_TAG", tag.c_str(), 1);
			execv("/proc/self/exe", argv);
			perror("execv");
			_exit(127);
		}
		kids.push_back(pid);
	}
	(void) argc;
	int worst = 0;
	for (pid_t k : kids) {
		int st = 0;
		waitpid(k, &st, 0);
		const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128;
		if (rc > worst) worst = rc;
	}
	return worst;
}

// rank 0 creates the NCCL id and publishes it (write + rename = atomic); the others wait for the file
static void ExchangeCommId(const std::string &dir, unsigned char id[128])
{
	const char *tag = getenv("METRO_RUN_TAG");
	if (!tag) tag = getenv("MASTER_PORT");
	const std::string path = dir + ".metro_nccl_id." + (tag ? tag : "0");
	if (locTask == 0) {
		if (metro_comm_unique_id(id) != METRO_OK) Fatal(std::string("cannot create the NCCL id: ") + metro_last_error(nullptr));
		const std::string tmp = path + ".tmp";
		FILE *f = fopen(tmp.c_str(), "wb");
		if (!f || fwrite(id, 1, 128, f) != 128) Fatal("cannot write " + tmp);
		fclose(f);
		if (rename(tmp.c_str(), path.c_str()) != 0) Fatal("cannot publish " + path);
	} else {
		for (int tries = 0;; tries++) {
			FILE *f = fopen(path.c_str(), "rb");
			if (f) {
				const size_t n = fread(id, 1, 128, f);
				fclose(f);
				if (n == 128) break;
			}
			if (tries > 6000) Fatal("rank 0 never published " + path);
			std::this_thread::sleep_for(std::chrono::milliseconds(10));
		}
	}
}

int main(int argc, char **argv)
{
	const char *usage = "usage: %s [--parse-only] [--ranks K] <config.cfg>\n";
	bool parseOnly = false;
	int ranksArg = 0;
	const char *cfgPath = nullptr;
	for (int a = 1; a < argc; a++) {
		if (!strcmp(argv[a], "--parse-only")) parseOnly = true;
		else if (!strcmp(argv[a], "--ranks") && a + 1 < argc) ranksArg = atoi(argv[++a]);
		else cfgPath = argv[a];
	}
	if (!cfgPath) { fprintf(stderr, usage, argv[0]); return 2; }
	static const char *const rankEnv[] = {"METRO_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", nullptr};
	static const char *const worldEnv[] = {"METRO_WORLD", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", nullptr};
	static const char *const localEnv[] = {"METRO_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", nullptr};
	if (ranksArg > 1 && !getenv("METRO_RANK")) return LaunchRanks(ranksArg, argc, argv);
	locTask = EnvInt(rankEnv, 0);
	totTask = EnvInt(worldEnv, 1);
	if (totTask < 1 || locTask < 0 || locTask >= totTask) { fprintf(stderr, "bad rank %d of %d\n", locTask, totTask); return 2; }
	gpuDevice = EnvInt(localEnv, locTask);
	IOSettings SettingsIO;
	SettingsIO.ReadConfigFile(cfgPath);
	InitLocVariables(!parseOnly);
	SettingsIO.Init();
	InitTrees(nSnapsUse);
	SettingsIO.DistributeFilesAmongTasks();
	if (!parseOnly && totTask > 1) {
		unsigned char id[128];
		ExchangeCommId(SettingsIO.pathOutput, id);
		if (metro_comm_init(gpuCtx, locTask, totTask, id) != METRO_OK) Fatal(std::string("NCCL communicator: ") + metro_last_error(gpuCtx));
	}
	if (parseOnly) {
		// host-side ingest check (no GPU): one line per catalogue with sizes and order-sensitive checksums
		for (iNumCat = 0; iNumCat < nSnapsUse; iNumCat++) {
			iUseCat = 0;
			SettingsIO.ReadHalos();
			SettingsIO.ReadParticles();
			unsigned long long ck = 1469598103934665603ull, hk = 1469598103934665603ull;
			const TupleSoA &T = locTuples[0];
			for (size_t i = 0; i < T.pid.size(); i++) ck = (ck ^ (T.pid[i] + 0x9E3779B97F4A7C15ull * (T.halo[i] + 1) + T.type[i])) * 1099511628211ull;
			for (const Halo &h : locHalos[0]) hk = (hk ^ (h.ID + (unsigned long long) h.nAllPart())) * 1099511628211ull;
			printf("PARSED snapshot=%s z=%.3f halos=%zu tuples=%zu halo_ck=%llu tuple_ck=%llu\n", SettingsIO.strSnaps[iNumCat].c_str(),
			       SettingsIO.redShift[iNumCat], locHalos[0].size(), T.pid.size(), hk, ck);
		}
		return 0;
	}
	if (locTask == 0)
		printf("\n\t\tMetroCPP-b200: merger trees on %d B200%s (%s, NPTYPES=%d%s, %s token rule)\n\n", totTask, totTask > 1 ? "s, tuples sharded by particle-ID range" : "",
		       metro_version(), nPTypes, noPType ? ", NOPTYPE" : "", treeFlavour == METRO_FLAVOUR_GATHER ? "GATHER_TREES" : "ZOOM");
	nTreeChunks = totTask;
	iUseCat = 0; iNumCat = 0;
	SettingsIO.ReadHalos();
	SettingsIO.ReadParticles();
	SettingsIO.WriteLog(0, 0.0f);
	if (locTask == 0) printf("Starting to loop on %d halo and particle files.\n", nSnapsUse);
	for (iNumCat = 1; iNumCat < nSnapsUse; iNumCat++) {
		double t0 = Now();
		iUseCat = 1;
		SettingsIO.ReadHalos();
		SettingsIO.ReadParticles();
		double t1 = Now();
		if (locTask == 0) printf("Files read in %gs.\n", t1 - t0);
		SettingsIO.WriteLog(iNumCat, (float) (t1 - t0));
		FindProgenitors(0, 1);                               // forwards (runs the device pipeline for the pair)
		FindProgenitors(1, 0);                               // backwards (same join, roles swapped)
		double t2 = Now();
		if (locTask == 0) printf("Finding halo progenitors, forwards and backwards... done in %gs.\n", t2 - t1);
		SettingsIO.WriteLog(iNumCat, (float) (t2 - t1));
		CleanTrees(iNumCat);
		SettingsIO.WriteTree(iNumCat);
		double t3 = Now();
		SettingsIO.WriteLog(iNumCat, (float) (t3 - t2));
		ShiftHalosPartsGrids();
		FreeMergerTrees(iNumCat);
		double t4 = Now();
		SettingsIO.WriteLog(iNumCat, (float) (t4 - t3));
		if (locTask == 0) printf("Memory cleared and Merger Trees written in %gs.\n", t4 - t2);
	}
	if (locTask == 0) printf("The loop on halo and particle catalogs has finished.\n");
	SettingsIO.WriteLog(-1, 0.0f);
	if (totTask > 1 && locTask == 0) {
		const char *tag = getenv("METRO_RUN_TAG");
		if (!tag) tag = getenv("MASTER_PORT");
		unlink((SettingsIO.pathOutput + ".metro_nccl_id." + (tag ? tag : "0")).c_str());
	}
	metro_ctx_destroy(gpuCtx);
	return 0;
}
