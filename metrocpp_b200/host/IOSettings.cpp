// IOSettings.cpp -- .cfg parser, AHF _halos/_particles readers producing the structure-of-arrays
// tuples the device ingests, and the .mtree writer.  File grammars follow the reference readers
// (reference src/IOSettings.cpp:97-213 cfg, :551-708 particles, :712-811 + :914-996 halos,
// :1074-1138 mtree, :1141-1173 timing log); the implementation is new: whole-file reads and a
// hand-rolled integer scanner instead of getline + sscanf + std::map inserts.
#include "IOSettings.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <map>
#include <memory>
#include <thread>
#include <unordered_map>

#include "global_vars.h"

static std::string Strip(const std::string &s)
{
	std::string o;
	for (char c : s) if (c != ' ' && c != '\t' && c != '\r' && c != '\n') o.push_back(c);
	return o;
}

// read-only view of a whole file (mmap: the page cache is parsed in place, no copy)
struct FileView {
	const char *p = nullptr;
	size_t n = 0;
	bool mapped = false;
	bool Open(const std::string &path)
	{
		int fd = open(path.c_str(), O_RDONLY);
		if (fd < 0) return false;
		struct stat st;
		if (fstat(fd, &st) != 0) { close(fd); return false; }
		n = (size_t) st.st_size;
		if (n == 0) { close(fd); p = ""; return true; }
		void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
		close(fd);
		if (m == MAP_FAILED) return false;
		madvise(m, n, MADV_SEQUENTIAL);
		p = (const char *) m; mapped = true;
		return true;
	}
	~FileView() { if (mapped) munmap((void *) p, n); }
};

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

static size_t CountLines(const char *p, const char *end)
{
	size_t n = 0;
	while (p < end) {
		const char *nl = (const char *) memchr(p, '\n', (size_t) (end - p));
		n++;
		if (!nl) break;
		p = nl + 1;
	}
	return n;
}

// The lines of an _halos file that start in [p, stop): every line is parsed on its own from its first character (a '#' in column 0
// is a comment, src/IOSettings.cpp:753), so chunks of whole lines can be parsed side by side.  `end` = end of the NUL-terminated buffer.
static void ParseHaloLines(const char *p, const char *stop, const char *end, int nPTypes, std::vector<Halo> &out)
{
	while (p < stop) {
		const char *q = SkipSpace(p, end);
		if (q >= end) break;
		if (*p == '#' || *q == '\n') { p = NextLine(p, end); continue; }
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
}

void IOSettings::ReadHalos()
{
	std::vector<Halo> &out = locHalos[iUseCat];
	out.clear();
	unsigned nThreads = std::thread::hardware_concurrency();
	const char *forced = getenv("METRO_INGEST_THREADS");               // tests: force the chunked path on small files
	if (forced && atoi(forced) > 0) nThreads = (unsigned) atoi(forced);
	if (nThreads == 0) nThreads = 1;
	if (nThreads > 64) nThreads = 64;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	for (const std::string &path : haloFiles[iNumCat]) {
		std::string buf;
		const double t0 = now();
		if (!ReadWholeFile(path, buf)) Fatal("file " + path + " not found");
		const double t1 = now();
		const char *base = buf.data(), *end = base + buf.size();
		size_t before = out.size();
		// chunks of whole lines, one per thread (10^6 halos are 100-400 MB of text: a second of strtof on one core)
		const unsigned nt = (buf.size() < (1u << 20) && !forced) ? 1u : nThreads;
		std::vector<const char *> cut(nt + 1, end);
		cut[0] = base;
		for (unsigned t = 1; t < nt; t++) {
			const char *c = base + buf.size() / nt * t;
			cut[t] = c <= cut[t - 1] ? cut[t - 1] : NextLine(c, end);
		}
		double t2 = t1;
		if (nt == 1) ParseHaloLines(base, end, end, nPTypes, out);
		else {
			std::vector<std::vector<Halo>> part(nt);
			std::vector<std::thread> th;
			for (unsigned t = 0; t < nt; t++)
				th.emplace_back([&, t] {
					part[t].reserve(CountLines(cut[t], cut[t + 1]));        // no regrowth while parsing
					ParseHaloLines(cut[t], cut[t + 1], end, nPTypes, part[t]);
				});
			for (auto &x : th) x.join();
			t2 = now();
			size_t total = out.size();
			for (auto &v : part) total += v.size();
			out.reserve(total);
			for (auto &v : part) out.insert(out.end(), v.begin(), v.end());
		}
		if (locTask == 0) printf("Reading %zu halos from file: %s\n", out.size() - before, path.c_str());
		if (getenv("METRO_DEBUG"))
			fprintf(stderr, "[metro] _halos ingest: %.1f MB, %zu halos, %u threads: read %.3f s, parse %.3f s, merge %.3f s\n", buf.size() / 1e6, out.size() - before, nt,
				t1 - t0, (nt == 1 ? now() : t2) - t1, nt == 1 ? 0.0 : now() - t2);
	}
	nLocHalos[iUseCat] = (int) out.size();
	slotDirty[iUseCat] = true;
}

// One text line of an _particles file, tokenised: first and second unsigned number (kNone = absent).
static const uint64_t kNone = ~0ull;

// Parses [p, end) (whole lines) into a[]/b[]: one entry per line, blank lines included.  Plain version: every access is bounds-checked
// (the last 64 bytes of a chunk, and anything the fast loop below hands back).
static size_t TokeniseTail(const char *p, const char *end, uint64_t *a, uint64_t *b)
{
	size_t k = 0;
	while (p < end) {
		const char *q = SkipSpace(p, end);
		uint64_t x = kNone, y = kNone;
		if (q < end && *q >= '0' && *q <= '9') {
			q = ScanU64(q, end, x);
			q = SkipSpace(q, end);
			if (q < end && *q >= '0' && *q <= '9') q = ScanU64(q, end, y);
		}
		a[k] = x; b[k] = y; k++;
		p = NextLine(q, end);
	}
	return k;
}

// Eight characters at a time: how many of the first bytes of w are decimal digits, and the value of eight digits (first character in
// the lowest byte; zero bytes count as the digit 0).  Same value modulo 2^64 as the digit-by-digit loop of ScanU64.
static inline uint64_t Load8(const char *p) { uint64_t w; memcpy(&w, p, 8); return w; }
static inline unsigned DigitRun(uint64_t w)
{
	const uint64_t d = w - 0x3030303030303030ull;
	const uint64_t nd = ((w + 0x4646464646464646ull) | d) & 0x8080808080808080ull;      // high bit set in every byte outside '0'..'9'
	return nd ? (unsigned) (__builtin_ctzll(nd) >> 3) : 8u;
}
static inline uint64_t Parse8(uint64_t w)
{
	w = (w & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;
	w = (w & 0x00FF00FF00FF00FFull) * 6553601 >> 16;
	return (w & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32;
}
static inline const char *ScanU64Fast(const char *p, const char *safe, uint64_t &v)      // 8 readable bytes behind every p < safe
{
	static const uint64_t pow10[8] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull};
	uint64_t x = 0;
	while (p < safe) {
		const uint64_t w = Load8(p);
		const unsigned n = DigitRun(w);
		if (n == 8) { x = x * 100000000ull + Parse8(w); p += 8; continue; }
		if (n) { x = x * pow10[n] + Parse8(w << ((8 - n) * 8)); p += n; }
		break;
	}
	v = x;
	return p;
}

// The same, without bounds checks while more than 64 bytes remain (172 MB, one thread: 0.53 -> 0.32 s).  A line that does not end
// inside that range is handed to the plain version from its first character.
static void TokeniseLines(const char *p, const char *end, uint64_t *a, uint64_t *b)
{
	size_t k = 0;
	const char *safe = end - p > 64 ? end - 64 : p;
	while (p < safe) {
		const char *q = p;
		while (q < safe && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
		if (q >= safe) break;
		uint64_t x = kNone, y = kNone;
		if ((unsigned) (*q - '0') <= 9u) {
			q = ScanU64Fast(q, safe, x);
			while (q < safe && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
			if (q >= safe) break;
			if ((unsigned) (*q - '0') <= 9u) q = ScanU64Fast(q, safe, y);
			if (q >= safe) break;
		}
		if (*q != '\n') {
			q = (const char *) memchr(q, '\n', (size_t) (safe - q));
			if (!q) break;
		}
		a[k] = x; b[k] = y; k++;
		p = q + 1;
	}
	TokeniseTail(p, end, a + k, b + k);
}

// AHF _particles reader (reference src/IOSettings.cpp:599-689: "nHalos", then per halo "nPart haloID" and
// nPart lines "pid [type]").  The grammar is sequential (a header is only recognisable by position), so the
// file is tokenised in parallel on all host cores (one entry per line), the header chain is walked serially
// (one step per halo), and the tuples are copied out in parallel.  Tuples come out grouped by halo; when the
// halos appear in catalogue order (the normal case) the per-halo offsets are kept for the CSR upload.
void IOSettings::ReadParticles()
{
	TupleSoA &T = locTuples[iUseCat];
	T.clear();
	const std::vector<Halo> &H = locHalos[iUseCat];
	std::unordered_map<uint64_t, uint32_t> byId;                       // only consulted when the order differs
	size_t iHalo = 0;
	size_t dropped = 0;
	bool inOrder = true;
	const bool verbose = getenv("METRO_DEBUG") != nullptr;
	auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tRead = 0, tTok = 0, tChain = 0, tCopy = 0;
	size_t bytes = 0;
	int64_t lastIdx = -1;
	struct Span { int64_t idx; size_t line, count, out; };
	unsigned nThreads = std::thread::hardware_concurrency();
	const char *forced = getenv("METRO_INGEST_THREADS");               // tests: force the chunked path on small files
	if (forced && atoi(forced) > 0) nThreads = (unsigned) atoi(forced);
	if (nThreads == 0) nThreads = 1;
	if (nThreads > 64) nThreads = 64;
	for (const std::string &path : partFiles[iNumCat]) {
		FileView buf;
		double t0 = now();
		if (!buf.Open(path)) Fatal("file " + path + " not found");
		if (locTask == 0) printf("Reading particle file: %s\n", path.c_str());
		const char *base = buf.p, *end = base + buf.n;
		bytes += buf.n;
		double t1 = now();
		tRead += t1 - t0;
		// chunks of whole lines
		const unsigned nt = (buf.n < (1u << 20) && !forced) ? 1u : nThreads;
		std::vector<const char *> cut(nt + 1, end);
		cut[0] = base;
		for (unsigned t = 1; t < nt; t++) {
			const char *c = base + buf.n / nt * t;
			cut[t] = c <= cut[t - 1] ? cut[t - 1] : NextLine(c, end);
		}
		std::vector<size_t> lineOff(nt + 1, 0);
		{
			std::vector<std::thread> th;
			for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t] { lineOff[t + 1] = CountLines(cut[t], cut[t + 1]); });
			for (auto &x : th) x.join();
		}
		for (unsigned t = 0; t < nt; t++) lineOff[t + 1] += lineOff[t];
		const size_t nLines = lineOff[nt];
		// token arrays: uninitialised (every entry is written below) and kept for the next file / snapshot - a fresh 16 bytes per line
		// would be page-faulted in again every time (a quarter of the tokenising time)
		static std::unique_ptr<uint64_t[]> tokA, tokB;
		static size_t tokCap = 0;
		if (nLines + 1 > tokCap) {
			tokA.reset(); tokB.reset();
			tokCap = nLines + 1 + nLines / 8;
			tokA.reset(new uint64_t[tokCap]); tokB.reset(new uint64_t[tokCap]);
		}
		uint64_t *const A = tokA.get(), *const B = tokB.get();
		{
			std::vector<std::thread> th;
			for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t] { TokeniseLines(cut[t], cut[t + 1], A + lineOff[t], B + lineOff[t]); });
			for (auto &x : th) x.join();
		}
		double t2 = now();
		tTok += t2 - t1;
		// header chain
		std::vector<Span> spans;
		size_t outPos = T.pid.size(), k = 0;
		while (k < nLines) {
			while (k < nLines && A[k] == kNone) k++;                   // blank lines between blocks
			if (k >= nLines) break;
			const uint64_t nFileHalos = A[k++];                        // line 1 of a block: halos in it (src/IOSettings.cpp:603-606)
			for (uint64_t fh = 0; fh < nFileHalos && k < nLines; fh++) {
				const uint64_t nPartHalo = A[k] == kNone ? 0 : A[k], haloID = B[k] == kNone ? 0 : B[k];   // "nPart haloID" (:610)
				k++;
				int64_t idx = -1;
				if (iHalo < H.size() && H[iHalo].ID == haloID) idx = (int64_t) iHalo++;
				else {
					if (byId.empty()) for (size_t h = 0; h < H.size(); h++) byId[H[h].ID] = (uint32_t) h;
					auto it = byId.find(haloID);
					if (it != byId.end()) idx = it->second;
				}
				const size_t cnt = (size_t) std::min<uint64_t>(nPartHalo, nLines - k);
				if (idx < 0) dropped += cnt;
				else {
					if (idx <= lastIdx) inOrder = false;
					lastIdx = idx;
					spans.push_back({idx, k, cnt, outPos});
					outPos += cnt;
				}
				k += cnt;
			}
		}
		double t3 = now();
		tChain += t3 - t2;
		// copy out
		T.pid.resize(outPos); T.halo.resize(outPos); T.type.resize(outPos);
		{
			std::vector<std::thread> th;
			// equal shares of the output range (halo sizes follow a power law: equal shares of spans would not balance)
			const size_t first = spans.empty() ? 0 : spans[0].out, total = outPos - first;
			for (unsigned t = 0; t < nt; t++)
				th.emplace_back([&, t] {
					const size_t lo = first + total * t / nt, hi = first + total * (t + 1) / nt;
					// first span that ends after lo
					size_t si = (size_t) (std::upper_bound(spans.begin(), spans.end(), lo, [](size_t v, const Span &sp) { return v < sp.out + sp.count; }) - spans.begin());
					for (; si < spans.size() && spans[si].out < hi; si++) {
						const Span &sp = spans[si];
						const size_t i0 = sp.out < lo ? lo - sp.out : 0, i1 = std::min(sp.count, hi - sp.out);
						for (size_t i = i0; i < i1; i++) {
							const uint64_t pid = A[sp.line + i], ty = B[sp.line + i];
							T.pid[sp.out + i] = pid == kNone ? 0 : pid;
							T.halo[sp.out + i] = (uint32_t) sp.idx;
							T.type[sp.out + i] = noPType ? 1 : (uint8_t) (ty == kNone ? 0 : std::min<uint64_t>(ty, 255));
						}
					}
				});
			for (auto &x : th) x.join();
		}
		if (!noPType)
			for (const Span &sp : spans)
				for (size_t i = 0; i < sp.count; i++)
					if (T.type[sp.out + i] >= nPTypes) Fatal("particle type >= nPTypes in " + path);
		if (inOrder) {
			if (T.haloOffset.empty()) T.haloOffset.assign(H.size() + 1, -1);
			for (const Span &sp : spans) T.haloOffset[sp.idx] = (int64_t) sp.out;
		}
		tCopy += now() - t3;
	}
	if (verbose)
		fprintf(stderr, "[metro] _particles ingest: %.1f MB, %zu tuples, %u threads: read %.3f s, tokenise %.3f s, header chain %.3f s, copy-out %.3f s\n",
			bytes / 1e6, T.pid.size(), nThreads, tRead, tTok, tChain, tCopy);
	if (inOrder && !H.empty()) {
		// halos without a block own an empty range; ranges are contiguous because spans were appended in order
		if (T.haloOffset.empty()) T.haloOffset.assign(H.size() + 1, -1);
		T.haloOffset[H.size()] = (int64_t) T.pid.size();
		for (size_t h = H.size(); h-- > 0;) if (T.haloOffset[h] < 0) T.haloOffset[h] = T.haloOffset[h + 1];
	} else {
		T.haloOffset.clear();
	}
	if (dropped) fprintf(stderr, "WARNING: %zu particle lines belong to halo IDs that are not in the halo catalogue; ignored\n", dropped);
	if (totTask > 1) ShardByParticleId(T, H.size());
	slotDirty[iUseCat] = true;
}

// Multi-rank runs: every rank reads ALL chunk files of the snapshot (so it holds the full halo table, with the catalogue
// npart the merit needs) and keeps the tuples of its particle-ID range: all memberships of a particle then sit on one GPU and the
// join needs no exchange.  The range width is fixed by the first (latest) snapshot: ceil(max ID / ranks) rounded up to a power of two
// share, the same on every rank and for every later snapshot (IDs above the last boundary stay with the last rank).
// (The reference hands nChunks / totTask chunk FILES to each task instead, src/IOSettings.cpp:459-460, and ships buffer halos between
// the tasks afterwards.)
void IOSettings::ShardByParticleId(TupleSoA &T, size_t nHalos)
{
	const size_t N = T.pid.size();
	if (pidShardWidth == 0) {
		uint64_t mx = 0;
		for (size_t i = 0; i < N; i++) mx = T.pid[i] > mx ? T.pid[i] : mx;
		uint64_t span = 1;
		while (span <= mx) span <<= 1;
		pidShardWidth = (span + (uint64_t) totTask - 1) / (uint64_t) totTask;
		if (pidShardWidth == 0) pidShardWidth = 1;
	}
	const uint64_t lo = pidShardWidth * (uint64_t) locTask;
	const bool last = locTask == totTask - 1;
	const uint64_t hi = lo + pidShardWidth;
	const bool grouped = T.haloOffset.size() == nHalos + 1;
	std::vector<int64_t> cnt(grouped ? nHalos + 1 : 0, 0);
	size_t w = 0;
	for (size_t i = 0; i < N; i++) {
		const uint64_t p = T.pid[i];
		if (p < lo || (!last && p >= hi)) continue;
		T.pid[w] = p; T.halo[w] = T.halo[i]; T.type[w] = T.type[i];
		if (grouped) cnt[T.halo[i] + 1]++;
		w++;
	}
	T.pid.resize(w); T.halo.resize(w); T.type.resize(w);
	if (grouped) {
		for (size_t h = 0; h < nHalos; h++) cnt[h + 1] += cnt[h];
		T.haloOffset = cnt;
	}
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
	if (locTask != 0) return;                                   // one timing log per run (the reference writes it on rank 0, src/main.cpp:156)
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
