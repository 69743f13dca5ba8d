// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include "ompl/base/goals/GoalState.h"
