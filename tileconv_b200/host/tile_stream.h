// tileconv_b200/host/tile_stream.h -- container drivers on either side of the hot path: read a TIS
// (with or without header) or a MOS / MOSC, push its tiles through the pool interface, write the
// TBC / MBC stream in tile order.  Follows the reference's drivers for the byte layout and the
// caller protocol (tileconv/graphics.cpp:66-169 tisToTBC, :278-392 mosToMBC, :773-973 readers,
// :1138-1165 writeEncodedTile; tileconv/FORMATS).
#ifndef TCB200_TILE_STREAM_H
#define TCB200_TILE_STREAM_H

#include <string>

#include "tc_iface.h"

namespace tc {

enum class InputKind { UNKNOWN, TIS, MOS };

InputKind sniffInput(const std::string& path, bool assumeTis) noexcept;
// Both return false on any error and remove the partial output (graphics.cpp:77,161).
bool encodeTisFile(const Options& options, const std::string& inFile, const std::string& outFile, std::string& error) noexcept;
bool encodeMosFile(const Options& options, const std::string& inFile, const std::string& outFile, std::string& error) noexcept;

// tileconv -I: prints the container facts of one file exactly like TileConv::showInfo (tileconv/tileconv.cpp:247-440)
// -- TIS / MOS / MOSC / TBC / MBC / TIZ / MOZ headers and the headerless-TIS guess under -T.  false where the
// reference returns false (unreadable, truncated or unknown files).
bool showInfo(const std::string& fileName, bool assumeTis) noexcept;

}  // namespace tc

#endif
