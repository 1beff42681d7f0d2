// stub: see tests/stubs/gtsam_stub.h
#pragma once
#include "../../gtsam_stub.h"
