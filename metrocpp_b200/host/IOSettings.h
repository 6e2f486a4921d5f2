// IOSettings.h -- configuration + AHF catalogue I/O, same method names as the reference's
// IOSettings (reference src/IOSettings.h:31-125).  The snapshot discovery that the reference
// delegates to scripts/find_*.sh + tmp/*.tmp caches (src/IOSettings.cpp:237-376) is done in-process.
#ifndef METRO_HOST_IOSETTINGS_H
#define METRO_HOST_IOSETTINGS_H
#include <fstream>
#include <string>
#include <vector>

class IOSettings {
public:
	std::string pathMetroCpp, pathInput, pathOutput, inputFormat;
	std::string haloSuffix, partSuffix, outSuffix, haloPrefix, partPrefix, outPrefix, cpuString, splitString;

	std::vector<float> redShift, aFactors;
	std::vector<std::string> strSnaps;                  // index 0 = LATEST snapshot (src/IOSettings.cpp:267-272)
	std::vector<int> numSnaps;
	std::vector<std::vector<std::string>> haloFiles, partFiles;   // [snapshot][chunk]

	void ReadConfigFile(const std::string &configFile);
	void Init();                                        // discover snapshots
	void DistributeFilesAmongTasks();                   // resolve file names per snapshot/chunk
	void ReadHalos();                                   // catalogue iNumCat -> locHalos[iUseCat]
	void ReadParticles();                               // catalogue iNumCat -> locTuples[iUseCat]
	void WriteTree(int iThisCat);                       // locCleanTrees[iThisCat-1] -> .mtree
	void WriteLog(int iNum, float seconds);

private:
	void InitFromCfgFile(const std::string &key, const std::string &value);
	void ShardByParticleId(struct TupleSoA &T, size_t nHalos);   // multi-rank: keep this rank's particle-ID range
	std::ofstream fileLogOut;
	std::vector<float> logTime;
};

#endif
