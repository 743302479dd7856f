// Integer tables of the reference LS estimator (product side, flat encoding).
//
//  * 256-tone VHT-LTF pattern with a single DC null:
//      packet_generation/phased_arr/helperMIMOChannelEstimate.m:16-23
//  * null / pilot carrier sets and CarriersLocations = setdiff(1:256, nulls U pilots):
//      packet_generation/phased_arr/generate_maMIMO_LTF.m:99-102
//
// tests/test_tables.py checks these bit for bit against (a) the oracle's structural
// re-derivation and (b) tests/golden/ref_tables.npz parsed from the reference source.
#pragma once
#include <stdint.h>

namespace mm {

// '+' = +1, '-' = -1, '0' = null tone; character i is MATLAB index i+1.
static const char kVhtLtf256[257] =
    "0000000++--++-+-++++++--++-+-++++++--++-+-+-----++--+-+-++++---+"
    "+-+-++-++--++-+-++++++--++-+-++++++--++-+-+-----++--+-+-+++++-+-"
    "0+--+++--++-+-++++++--++-+-++++++--++-+-+-----++--+-+-++++---++-"
    "+-++-++--++-+-++++++--++-+-++++++--++-+-+-----++--+-+-++++000000";

static const int kFftLen = 256;
// 1-based, generate_maMIMO_LTF.m:100
static const int kPilotCarriers[8] = {26, 54, 90, 118, 140, 168, 204, 232};

inline bool is_null_carrier(int idx1) {      // generate_maMIMO_LTF.m:99  [1:7 129 256-5:256]
  return idx1 <= 7 || idx1 == 129 || idx1 >= 251;
}
inline bool is_pilot_carrier(int idx1) {
  for (int i = 0; i < 8; ++i)
    if (kPilotCarriers[i] == idx1) return true;
  return false;
}

}  // namespace mm
