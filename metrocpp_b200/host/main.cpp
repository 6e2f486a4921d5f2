// main.cpp -- MetroCPP-b200 <config.cfg>: the reference's driver loop (reference src/main.cpp:48-328)
// with the matching path on the GPU.  Snapshot 0 is the latest catalogue; the loop walks backwards
// in time: read the older catalogue, match the pair, clean, write the .mtree of the later snapshot,
// shift 1 -> 0.
#include <chrono>
#include <cstdio>
#include <string>

#include "IOSettings.h"
#include "MergerTree.h"
#include "global_vars.h"

static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: %s [--parse-only] <config.cfg>\n", argv[0]); return 2; }
	const bool parseOnly = std::string(argv[1]) == "--parse-only";
	if (parseOnly && argc < 3) { fprintf(stderr, "usage: %s --parse-only <config.cfg>\n", argv[0]); return 2; }
	IOSettings SettingsIO;
	SettingsIO.ReadConfigFile(argv[parseOnly ? 2 : 1]);
	InitLocVariables(!parseOnly);
	SettingsIO.Init();
	InitTrees(nSnapsUse);
	SettingsIO.DistributeFilesAmongTasks();
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
	printf("\n\t\tMetroCPP-b200: merger trees on one B200 (%s, NPTYPES=%d%s, %s token rule)\n\n", metro_version(), nPTypes,
	       noPType ? ", NOPTYPE" : "", treeFlavour == METRO_FLAVOUR_GATHER ? "GATHER_TREES" : "ZOOM");
	nTreeChunks = totTask;
	iUseCat = 0; iNumCat = 0;
	SettingsIO.ReadHalos();
	SettingsIO.ReadParticles();
	SettingsIO.WriteLog(0, 0.0f);
	printf("Starting to loop on %d halo and particle files.\n", nSnapsUse);
	for (iNumCat = 1; iNumCat < nSnapsUse; iNumCat++) {
		double t0 = Now();
		iUseCat = 1;
		SettingsIO.ReadHalos();
		SettingsIO.ReadParticles();
		double t1 = Now();
		printf("Files read in %gs.\n", t1 - t0);
		SettingsIO.WriteLog(iNumCat, (float) (t1 - t0));
		FindProgenitors(0, 1);                               // forwards (runs the device pipeline for the pair)
		FindProgenitors(1, 0);                               // backwards (same join, roles swapped)
		double t2 = Now();
		printf("Finding halo progenitors, forwards and backwards... done in %gs.\n", t2 - t1);
		SettingsIO.WriteLog(iNumCat, (float) (t2 - t1));
		CleanTrees(iNumCat);
		SettingsIO.WriteTree(iNumCat);
		double t3 = Now();
		SettingsIO.WriteLog(iNumCat, (float) (t3 - t2));
		ShiftHalosPartsGrids();
		FreeMergerTrees(iNumCat);
		double t4 = Now();
		SettingsIO.WriteLog(iNumCat, (float) (t4 - t3));
		printf("Memory cleared and Merger Trees written in %gs.\n", t4 - t2);
	}
	printf("The loop on halo and particle catalogs has finished.\n");
	SettingsIO.WriteLog(-1, 0.0f);
	metro_ctx_destroy(gpuCtx);
	return 0;
}
