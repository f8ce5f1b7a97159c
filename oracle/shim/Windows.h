/* Stand-in for <Windows.h> so the reference's zip loader compiles with g++ on Linux.
 * Test infrastructure only (used by oracle/build_ref.sh); not part of the product. */
#pragma once
#include <strings.h>
#include <stdint.h>
#ifndef _stricmp
#define _stricmp strcasecmp
#endif
