// IOSettings.cpp -- .cfg parser, AHF _halos/_particles readers producing the structure-of-arrays
// tuples the device ingests, and the .mtree writer.  File grammars follow the reference readers
// (reference src/IOSettings.cpp:97-213 cfg, :551-708 particles, :712-811 + :914-996 halos,
// :1074-1138 mtree, :1141-1173 timing log); the implementation is new: whole-file reads and a
// hand-rolled integer scanner instead of getline + sscanf + std::map inserts.
#include "IOSettings.h"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <map>
#include <unordered_map>

#include "global_vars.h"

static std::string Strip(const std::string &s)
{
	std::string o;
	for (char c : s) if (c != ' ' && c != '\t' && c != '\r' && c != '\n') o.push_back(c);
	return o;
}

static bool ReadWholeFile(const std::string &path, std::string &out)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	out.resize((size_t) n);
	size_t got = n > 0 ? fread(&out[0], 1, (size_t) n, f) : 0;
	fclose(f);
	return got == (size_t) n;
}

void IOSettings::ReadConfigFile(const std::string &configFile)
{
	std::ifstream in(configFile);
	if (!in.good()) Fatal("config file " + configFile + " not found");
	std::string line;
	while (std::getline(in, line)) {
		if (line.empty() || line[0] == '#' || line[0] == ' ' || line[0] == '\t') continue;   // src/IOSettings.cpp:125
		size_t eq = line.find('=');
		if (eq == std::string::npos) continue;
		std::string key = Strip(line.substr(0, eq)), val = Strip(line.substr(eq + 1));
		if (!key.empty()) InitFromCfgFile(key, val);
	}
}

void IOSettings::InitFromCfgFile(const std::string &k, const std::string &v)
{
	auto toi = [&](const char *what) { if (v.empty()) Fatal(std::string(what) + " has no value"); return atoi(v.c_str()); };
	if (k == "boxSize") boxSize = (float) atof(v.c_str());
	else if (k == "haloSuffix") haloSuffix = v;
	else if (k == "partSuffix") partSuffix = v;
	else if (k == "outSuffix") outSuffix = v;
	else if (k == "haloPrefix") haloPrefix = v;
	else if (k == "partPrefix") partPrefix = v;
	else if (k == "outPrefix") outPrefix = v;
	else if (k == "cpuString") cpuString = v;
	else if (k == "splitString") splitString = v;
	else if (k == "pathMetroCpp") pathMetroCpp = v;
	else if (k == "pathInput") pathInput = v;
	else if (k == "inputFormat") inputFormat = v;
	else if (k == "nSnapsUse") nSnapsUse = toi("nSnapsUse");
	else if (k == "nChunks") nChunks = toi("nChunks");
	else if (k == "nGrid") nGrid = toi("nGrid");
	else if (k == "facOrphanSteps") facOrphanSteps = toi("facOrphanSteps");
	else if (k == "maxOrphanSteps") maxOrphanSteps = toi("maxOrphanSteps");
	else if (k == "minPartHalo") minPartHalo = toi("minPartHalo");
	else if (k == "minPartCmp") minPartCmp = toi("minPartCmp");
	else if (k == "pathOutput") pathOutput = v;
	else if (k == "nTreeChunks") nTreeChunks = toi("nTreeChunks");
	else if (k == "cosmologicalModel") cosmologicalModel = v;
	// ---- run-time stand-ins for the reference's compile-time switches (Makefile.config:29-50)
	else if (k == "nPTypes") nPTypes = toi("nPTypes");                   // -DNPTYPES=N
	else if (k == "noPType") noPType = toi("noPType") != 0;              // -DNOPTYPE
	else if (k == "treeFlavour") {                                       // -DZOOM | -DGATHER_TREES
		if (v == "zoom") treeFlavour = METRO_FLAVOUR_ZOOM_DUP;
		else if (v == "gather") treeFlavour = METRO_FLAVOUR_GATHER;
		else Fatal("treeFlavour must be zoom or gather");
	} else if (k == "keepRawTrees") keepRawTrees = toi("keepRawTrees") != 0;
	else if (k == "runMode" || k == "pathTree") { /* parsed by nobody in the reference either */ }
	else printf("Arg= %s is useless or redundant and will be ignored.\n", k.c_str());
	if (k == "nChunks" && nChunks < 1) Fatal("nChunks must be a positive integer");
	if (k == "nPTypes" && (nPTypes < 1 || nPTypes > NPTYPES_MAX)) Fatal("nPTypes out of range");
}

// anyname_NNN[.CCCC].zZ.ZZZ.suffix  ->  (NNN, chunk or -1, "Z.ZZZ")
static bool ParseAhfName(const std::string &name, const std::string &prefix, const std::string &suffix, std::string &num, int &chunk,
			 std::string &z)
{
	if (name.size() <= prefix.size() + suffix.size() || name.compare(0, prefix.size(), prefix) != 0) return false;
	if (name.compare(name.size() - suffix.size(), suffix.size(), suffix) != 0) return false;
	std::string mid = name.substr(prefix.size(), name.size() - prefix.size() - suffix.size());   // NNN[.CCCC].zZ.ZZZ.
	if (mid.size() < 4 || mid.back() != '.') return false;
	mid.pop_back();
	size_t zp = mid.find(".z");
	if (zp == std::string::npos) return false;
	z = mid.substr(zp + 2);
	std::string head = mid.substr(0, zp);                 // NNN or NNN.CCCC
	size_t dot = head.find('.');
	num = head.substr(0, dot);
	chunk = dot == std::string::npos ? -1 : atoi(head.substr(dot + 1).c_str());
	if (num.empty() || num.find_first_not_of("0123456789") != std::string::npos) return false;
	return !z.empty();
}

void IOSettings::Init()
{
	if (inputFormat != "AHF") Fatal("input format '" + inputFormat + "' is not supported (AHF only)");
	if (!pathInput.empty() && pathInput.back() != '/') pathInput += '/';
	if (!pathOutput.empty() && pathOutput.back() != '/') pathOutput += '/';
	DIR *d = opendir(pathInput.c_str());
	if (!d) Fatal("pathInput " + pathInput + " cannot be opened");
	std::map<int, std::pair<std::string, std::string>> snaps;      // number -> (string NNN, redshift string)
	while (dirent *e = readdir(d)) {
		std::string num, z;
		int chunk;
		if (ParseAhfName(e->d_name, haloPrefix, haloSuffix, num, chunk, z)) snaps[atoi(num.c_str())] = {num, z};
	}
	closedir(d);
	nSnaps = (int) snaps.size();
	if (nSnaps < 2) Fatal("fewer than two AHF_halos catalogues found in " + pathInput);
	if (nSnapsUse <= 0 || nSnapsUse > nSnaps) nSnapsUse = nSnaps;
	strSnaps.assign(nSnaps, ""); numSnaps.assign(nSnaps, 0); redShift.assign(nSnaps, 0.f); aFactors.assign(nSnaps, 0.f);
	int i = nSnaps;
	for (auto &kv : snaps) {                                   // ascending number -> filled back to front
		--i;
		numSnaps[i] = kv.first;
		strSnaps[i] = kv.second.first;
		redShift[i] = (float) atof(kv.second.second.c_str());
		aFactors[i] = 1.0f / (1.0f + redShift[i]);
	}
	if (locTask == 0) printf("Found %d redshifts in total. Using %s file format.\n", nSnaps, inputFormat.c_str());
}

void IOSettings::DistributeFilesAmongTasks()
{
	haloFiles.assign(nSnaps, {});
	partFiles.assign(nSnaps, {});
	DIR *d = opendir(pathInput.c_str());
	if (!d) Fatal("pathInput " + pathInput + " cannot be opened");
	std::map<std::pair<int, int>, std::string> halo, part;       // (number, chunk) -> file name
	while (dirent *e = readdir(d)) {
		std::string num, z;
		int chunk;
		if (ParseAhfName(e->d_name, haloPrefix, haloSuffix, num, chunk, z)) halo[{atoi(num.c_str()), chunk}] = e->d_name;
		if (ParseAhfName(e->d_name, partPrefix, partSuffix, num, chunk, z)) part[{atoi(num.c_str()), chunk}] = e->d_name;
	}
	closedir(d);
	for (int s = 0; s < nSnaps; s++) {
		for (auto &kv : halo) {
			if (kv.first.first != numSnaps[s]) continue;
			auto it = part.find(kv.first);
			if (it == part.end()) Fatal("no " + partSuffix + " file next to " + kv.second);
			haloFiles[s].push_back(pathInput + kv.second);
			partFiles[s].push_back(pathInput + it->second);
		}
		if (haloFiles[s].empty()) Fatal("no catalogue files for snapshot " + strSnaps[s]);
	}
}

static inline const char *SkipSpace(const char *p, const char *end) { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++; return p; }
static inline const char *NextLine(const char *p, const char *end) { while (p < end && *p != '\n') p++; return p < end ? p + 1 : end; }
static inline const char *ScanU64(const char *p, const char *end, uint64_t &v)
{
	p = SkipSpace(p, end);
	uint64_t x = 0;
	while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (uint64_t) (*p++ - '0');
	v = x;
	return p;
}

void IOSettings::ReadHalos()
{
	std::vector<Halo> &out = locHalos[iUseCat];
	out.clear();
	for (const std::string &path : haloFiles[iNumCat]) {
		std::string buf;
		if (!ReadWholeFile(path, buf)) Fatal("file " + path + " not found");
		const char *p = buf.data(), *end = p + buf.size();
		size_t before = out.size();
		while (p < end) {
			const char *q = SkipSpace(p, end);
			if (q >= end) break;
			if (*p == '#' || *q == '\n') { p = NextLine(p, end); continue; }        // '#' in column 0 (src/IOSettings.cpp:753)
			Halo h;
			char *e;
			h.ID = strtoull(q, &e, 10);
			(void) strtol(e, &e, 10);                                  // hostHalo: read into a dummy by the reference
			h.nSub = (int) strtol(e, &e, 10);
			h.mTot = strtof(e, &e);
			long npart = strtol(e, &e, 10);
			for (int i = 0; i < 3; i++) h.X[i] = strtof(e, &e);
			for (int i = 0; i < 3; i++) h.V[i] = strtof(e, &e);
			// ReadLineAHF (src/IOSettings.cpp:961-989): all particles are booked under type 1
			h.nPart[0] = 0;
			if (nPTypes > 1) h.nPart[1] = (int) npart; else h.nPart[0] = (int) npart;
			out.push_back(h);
			p = NextLine(p, end);
		}
		if (locTask == 0) printf("Reading %zu halos from file: %s\n", out.size() - before, path.c_str());
	}
	nLocHalos[iUseCat] = (int) out.size();
	slotDirty[iUseCat] = true;
}

void IOSettings::ReadParticles()
{
	TupleSoA &T = locTuples[iUseCat];
	T.clear();
	const std::vector<Halo> &H = locHalos[iUseCat];
	std::unordered_map<uint64_t, uint32_t> byId;                       // only consulted when the order differs
	size_t iHalo = 0;
	size_t dropped = 0;
	for (const std::string &path : partFiles[iNumCat]) {
		std::string buf;
		if (!ReadWholeFile(path, buf)) Fatal("file " + path + " not found");
		if (locTask == 0) printf("Reading particle file: %s\n", path.c_str());
		const char *p = buf.data(), *end = p + buf.size();
		T.pid.reserve(T.pid.size() + buf.size() / 12);
		while (p < end) {
			uint64_t nFileHalos = 0;
			p = NextLine(ScanU64(p, end, nFileHalos), end);            // line 1: halos in this block (src/IOSettings.cpp:603-606)
			for (uint64_t fh = 0; fh < nFileHalos && p < end; fh++) {
				uint64_t nPartHalo = 0, haloID = 0;
				p = ScanU64(p, end, nPartHalo);
				p = NextLine(ScanU64(p, end, haloID), end);            // "nPart haloID" (:610)
				int64_t idx = -1;
				if (iHalo < H.size() && H[iHalo].ID == haloID) idx = (int64_t) iHalo++;
				else {
					if (byId.empty()) for (size_t h = 0; h < H.size(); h++) byId[H[h].ID] = (uint32_t) h;
					auto it = byId.find(haloID);
					if (it != byId.end()) idx = it->second;
				}
				for (uint64_t k = 0; k < nPartHalo && p < end; k++) {
					uint64_t pid = 0, type = 1;
					p = ScanU64(p, end, pid);
					if (!noPType) { p = ScanU64(p, end, type); if (type >= (uint64_t) nPTypes) Fatal("particle type >= nPTypes in " + path); }
					p = NextLine(p, end);
					if (idx < 0) { dropped++; continue; }
					T.pid.push_back(pid);
					T.halo.push_back((uint32_t) idx);
					T.type.push_back((uint8_t) type);
				}
			}
			p = SkipSpace(p, end);
			while (p < end && *p == '\n') p = SkipSpace(p + 1, end);
		}
	}
	if (dropped) fprintf(stderr, "WARNING: %zu particle lines belong to halo IDs that are not in the halo catalogue; ignored\n", dropped);
	slotDirty[iUseCat] = true;
}

void IOSettings::WriteTree(int iThisCat)
{
	const int iC = iThisCat - 1;
	const std::string outName = pathOutput + outPrefix + strSnaps[iC] + "." + std::to_string(locTask) + "." + outSuffix;
	FILE *f = fopen(outName.c_str(), "w");
	if (!f) Fatal("cannot write " + outName);
	if (locTask == 0) {
		printf("Printing trees to file %s\n", outName.c_str());
		fputs("# ID host(1)   N particles host(2)   Num. progenitors(3)  Orphan[0=no, 1=yes](4)\n", f);
		fputs("# Total particles (1)   ID progenitor(2)   Particles in common (3)\n", f);
	}
	for (const MergerTree &t : locCleanTrees[iC]) {
		fprintf(f, "%llu %d %zu %d\n", (unsigned long long) t.mainHalo.ID, t.mainHalo.nAllPart(), t.idProgenitor.size(), t.isOrphan ? 1 : 0);
		for (size_t p = 0; p < t.idProgenitor.size(); p++) {
			int comm = 0;
			for (int ty = 0; ty < nPTypes; ty++) comm += t.nCommon[ty][p];
			fprintf(f, "%d %llu %d\n", t.progHalo[p].nAllPart(), (unsigned long long) t.progHalo[p].ID, comm);
		}
	}
	fclose(f);
}

void IOSettings::WriteLog(int iNum, float seconds)
{
	if (iNum == 0) {
		const std::string name = pathOutput + "timing_log." + std::to_string(totTask) + "." + std::to_string(nChunks) + ".b200.txt";
		fileLogOut.open(name);
		fileLogOut << "# ReadFile (1) ForwardTree+BackwardTree (2) CleanTrees+Write (3) Shift (4)" << std::endl;
	} else if (iNum > 0) {
		logTime.push_back(seconds);
		if (logTime.size() == 4) {
			for (float t : logTime) fileLogOut << std::setw(9) << t << "\t";
			fileLogOut << std::endl;
			logTime.clear();
		}
	} else {
		fileLogOut.close();
	}
}
