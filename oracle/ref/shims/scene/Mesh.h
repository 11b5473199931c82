#pragma once
#include "../inc/Mesh.h"
