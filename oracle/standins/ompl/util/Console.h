// OMPL API stand-in (oracle/standins/README.md): logging is dropped.
#pragma once
#define OMPL_INFORM(...) ((void)0)
#define OMPL_WARN(...) ((void)0)
#define OMPL_ERROR(...) ((void)0)
#define OMPL_DEBUG(...) ((void)0)
