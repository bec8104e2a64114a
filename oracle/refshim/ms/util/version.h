// Stand-in for the header the reference's build system would generate from
// libs/ms/inc/ms/util/version.h.in (cmake configure_file). Written for the oracle
// build recipe (oracle/Makefile); only provides the version string macro + externs.
#pragma once
#include <string>
#include "util/support.h"
#define MA_VERSION "ref-oracle-build"
extern DLL_PORT(MS) const std::string sLibMaVersion;
extern DLL_PORT(MS) const bool bLibMaWithPython;
