// MergerTree.cpp -- FindProgenitors / CleanTrees / ShiftHalosPartsGrids on top of the C ABI.
// Same names, argument meaning and call order as the reference (src/MergerTree.cpp:399-652,
// src/utils.cpp:245-403); the work happens in libmetro_b200.so.
#include "MergerTree.h"

#include <cstdio>
#include <string>

#include "global_vars.h"

void MergerTree::Clean()
{
	progHalo.clear(); progHalo.shrink_to_fit();
	idProgenitor.clear(); idProgenitor.shrink_to_fit();
	nCommon.clear(); nCommon.shrink_to_fit();
}

void InitTrees(int nUseCat) { locCleanTrees.assign(nUseCat > 1 ? nUseCat - 1 : 0, {}); }

static metro_result_info lastInfo;
static bool haveResult = false;

static void Check(int rc, const char *what)
{
	if (rc != METRO_OK) Fatal(std::string(what) + ": " + metro_last_error(gpuCtx));
}

static void UploadSlot(int slot)
{
	if (!slotDirty[slot]) return;
	const std::vector<Halo> &H = locHalos[slot];
	const size_t n = H.size();
	std::vector<uint64_t> id(n);
	std::vector<int32_t> np(n * nPTypes), orph(n);
	std::vector<uint8_t> tok(n);
	for (size_t h = 0; h < n; h++) {
		id[h] = H[h].ID;
		orph[h] = H[h].nOrphanSteps;
		tok[h] = H[h].isToken ? 1 : 0;
		for (int t = 0; t < nPTypes; t++) np[h * nPTypes + t] = H[h].nPart[t];
	}
	const TupleSoA &T = locTuples[slot];
	metro_catalogue c;
	c.n_halos = (int64_t) n;
	c.halo_id = id.data(); c.npart = np.data(); c.n_orphan_steps = orph.data(); c.is_token = tok.data();
	c.n_tuples = (int64_t) T.pid.size();
	c.pid = T.pid.data(); c.halo = T.halo.data();
	c.type = noPType ? nullptr : T.type.data();
	metro_params prm = CurrentParams();
	// tuples grouped by halo in catalogue order (the normal _particles layout): 8 instead of 12 bytes per tuple go up
	if (T.haloOffset.size() == n + 1 && n > 0) Check(metro_catalogue_upload_csr(gpuCtx, slot, &prm, &c, T.haloOffset.data()), "catalogue upload");
	else Check(metro_catalogue_upload(gpuCtx, slot, &prm, &c), "catalogue upload");
	slotDirty[slot] = false;
}

// raw tree `dir` of the last match -> locMTrees[dir]
static void MaterialiseRaw(int dir)
{
	const int other = 1 - dir;
	const int64_t H = lastInfo.n_halos[dir], M = lastInfo.raw_count[dir];
	std::vector<int64_t> off(H + 1);
	std::vector<int32_t> prog(M), nc((size_t) M * nPTypes);
	metro_result_arrays a = {};
	a.raw_off[dir] = off.data(); a.raw_prog[dir] = prog.data(); a.raw_ncommon[dir] = nc.data();
	Check(metro_result_fetch(gpuCtx, &a), "result fetch");
	std::vector<MergerTree> &out = locMTrees[dir];
	out.assign(H, MergerTree());
	for (int64_t h = 0; h < H; h++) {
		MergerTree &t = out[h];
		t.mainHalo = locHalos[dir][h];
		t.nCommon.assign(nPTypes, {});
		const int64_t lo = off[h], hi = off[h + 1];
		t.isOrphan = lo == hi;                       // MergerTree::AssignMap (src/MergerTree.cpp:157-168)
		for (int64_t c = lo; c < hi; c++) {
			const Halo &p = locHalos[other][prog[c]];
			t.progHalo.push_back(p);
			t.idProgenitor.push_back(p.ID);
			for (int ty = 0; ty < nPTypes; ty++) t.nCommon[ty].push_back(nc[c * nPTypes + ty]);
		}
	}
}

void FindProgenitors(int iOne, int iTwo)
{
	if (iOne == iTwo || iOne < 0 || iOne > 1 || iTwo < 0 || iTwo > 1) Fatal("FindProgenitors: catalogues must be (0,1) or (1,0)");
	if (iOne == 0) {
		UploadSlot(0);
		UploadSlot(1);
		metro_params prm = CurrentParams();
		Check(metro_match_pair(gpuCtx, &prm, &lastInfo), "progenitor matching");
		haveResult = true;
	} else if (!haveResult) {
		Fatal("FindProgenitors(1,0) must follow FindProgenitors(0,1) for the same pair");
	}
	if (keepRawTrees) MaterialiseRaw(iOne);     // locMTrees[iOne]; main() only needs the clean trees
}

void CleanTrees(int iStep)
{
	if (!haveResult) Fatal("CleanTrees before FindProgenitors");
	if (iStep < 1 || iStep > (int) locCleanTrees.size()) Fatal("CleanTrees: step out of range");
	const int64_t nt = lastInfo.n_trees, cc = lastInfo.clean_count, nk = lastInfo.n_tokens;
	std::vector<int32_t> main(nt), prog(cc), nc((size_t) cc * nPTypes), tok(nk);
	std::vector<uint8_t> orph(nt);
	std::vector<int64_t> off(nt + 1);
	metro_result_arrays a = {};
	a.tree_main = main.data(); a.tree_orphan = orph.data(); a.tree_off = off.data();
	a.clean_prog = prog.data(); a.clean_ncommon = nc.data(); a.token_halo = tok.data();
	Check(metro_result_fetch(gpuCtx, &a), "result fetch");
	std::vector<MergerTree> &out = locCleanTrees[iStep - 1];
	out.clear();
	out.reserve(nt);
	for (int64_t k = 0; k < nt; k++) {
		MergerTree t;
		t.mainHalo = locHalos[0][main[k]];
		t.isOrphan = orph[k] != 0;
		t.nCommon.assign(nPTypes, {});
		for (int64_t c = off[k]; c < off[k + 1]; c++) {
			Halo p;
			if (prog[c] < 0) {                       // token copy of the orphan itself (src/MergerTree.cpp:577-579)
				p = t.mainHalo;
				p.nOrphanSteps++;
				p.isToken = true;
			} else {
				p = locHalos[1][prog[c]];
			}
			t.progHalo.push_back(p);
			t.idProgenitor.push_back(p.ID);
			for (int ty = 0; ty < nPTypes; ty++) t.nCommon[ty].push_back(nc[c * nPTypes + ty]);
		}
		out.push_back(std::move(t));
	}
	allOrphIDs.clear();
	for (int64_t k = 0; k < nk; k++) allOrphIDs.push_back(locHalos[0][tok[k]].ID);
	if (locTask == 0)
		printf("The total number of orphan halos after cleaning the connections is: %lld while %lld will be untracked.\n",
		       (long long) nk, (long long) lastInfo.n_untracked);
}

void FreeMergerTrees(int iStep)
{
	locMTrees[0].clear(); locMTrees[0].shrink_to_fit();
	locMTrees[1].clear(); locMTrees[1].shrink_to_fit();
	if (iStep >= 1 && iStep <= (int) locCleanTrees.size()) { locCleanTrees[iStep - 1].clear(); locCleanTrees[iStep - 1].shrink_to_fit(); }
}

void ShiftHalosPartsGrids()
{
	if (!haveResult) Fatal("ShiftHalosPartsGrids before FindProgenitors/CleanTrees");
	// token halos: copies of the orphans with nOrphanSteps+1, appended after the real halos
	std::vector<int32_t> tok(lastInfo.n_tokens);
	metro_result_arrays a = {};
	a.token_halo = tok.data();
	Check(metro_result_fetch(gpuCtx, &a), "result fetch");
	std::vector<Halo> next = locHalos[1];
	for (int32_t h : tok) {
		Halo t = locHalos[0][h];
		t.nOrphanSteps++;
		t.isToken = true;
		next.push_back(t);
	}
	metro_params prm = CurrentParams();
	Check(metro_shift(gpuCtx, &prm), "snapshot shift");      // tuples move on the device
	locHalos[0].swap(next);
	nLocHalos[0] = (int) locHalos[0].size();
	locTuples[0].clear();                                     // resident on the device only
	slotDirty[0] = false;
	CleanMemory(1);
	haveResult = false;
}
