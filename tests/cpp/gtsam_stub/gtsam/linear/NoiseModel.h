// TEST HARNESS ONLY: see gtsam_stub_core.h
#pragma once
#include "../../gtsam_stub_core.h"
