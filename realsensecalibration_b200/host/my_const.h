// Compile-time configuration of the reference (Main_Calibration/my_const.h:9-16), same names and values; they
// parameterise the committed golden run.  The B200 host classes read them as defaults only: every one of them can
// be overridden at run time (BAManager::Configure), because the C ABI takes them as arguments.
#pragma once
#include <string>

namespace RSCalibration {
const static double MARKER_SIDE = 0.0148;
const static int TIMES = 6;
const static int CAMERAS = 4;
const static int MARKERS = 11;
const static int BASE_MARKER_ID = 0;
const static std::string SERIAL_NUMBERS[4] = {"821312061029", "816612062327", "821212062536", "821212061326"};
const static int MARKER_IDS[11] = {0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 23};
}  // namespace RSCalibration
