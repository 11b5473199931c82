#pragma once
// CMake-generated in the reference (utils/RootDir.h.in:2)
#define ROOT_DIR REF_ROOT
