// nw_registry_b200.cpp -- getNwAlgorithmMap with the B200 entry added.
//
// The reference's nw_algorithm.cpp is compiled UNMODIFIED with -DgetNwAlgorithmMap=getNwAlgorithmMap_reference
// (the reference build recipe); this translation unit provides the symbol benchmark.cpp and cmd_parser.cpp link against and
// appends one line to the registry -- the whole integration a maintainer needs.
#include "nw_algorithm.hpp"
#include "nw_fns.hpp"

NwStat NwAlign_B200(const NwAlgParams& pr, NwAlgInput& nw, NwAlgResult& res);
NwStat NwTrace_B200(NwAlgInput& nw, NwAlgResult& res, bool calcDebugTrace);
NwStat NwHash_B200(NwAlgInput& nw, NwAlgResult& res);
NwStat NwPrintScore_B200(std::ostream& os, const NwAlgInput& nw, NwAlgResult& res);

void getNwAlgorithmMap_reference(Dict<std::string, NwAlgorithm>& algMap);

void getNwAlgorithmMap(Dict<std::string, NwAlgorithm>& algMap)
{
    getNwAlgorithmMap_reference(algMap);
    algMap.insert(std::string("NwAlign_B200"), NwAlgorithm {NwAlign_B200, NwTrace_B200, NwHash_B200, NwPrintScore_B200, NwPrintTrace1_Plain});
}
