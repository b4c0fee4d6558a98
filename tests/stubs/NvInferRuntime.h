// stub: see NvInfer.h in this directory
#pragma once
#include "NvInfer.h"
