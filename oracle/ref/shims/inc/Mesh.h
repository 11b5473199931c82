#pragma once
// the reference spells this header "Mesh.h" but ships it as "mesh.h" (case-insensitive macOS FS)
#include <mesh.h>
