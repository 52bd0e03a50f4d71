// Selects where the reference-facing declarations (tc::Converter, tc::Options, tc::TileData,
// tc::TileThreadPool, ...) come from: the reference's own headers when the plugin is built INTO
// tileconv (-DTCB200_WITH_REFERENCE_HEADERS -I<reference>/tileconv), or this
// repo's stand-alone mirror otherwise.
#ifndef TCB200_TC_IFACE_H
#define TCB200_TC_IFACE_H
#ifdef TCB200_WITH_REFERENCE_HEADERS
#include "types.h"
#include "options.h"
#include "converter.h"
#include "converter_dxt.h"
#include "converterfactory.h"
#include "tiledata.h"
#include "tilethreadpool_base.h"
#else
#include "tc_mirror.h"
#endif
#endif
