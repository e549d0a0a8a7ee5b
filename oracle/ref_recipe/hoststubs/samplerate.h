// TEST INFRASTRUCTURE ONLY. Stands in for libsamplerate's header so that the reference's
// waveguide/src/config.cpp compiles whole; only its three closed-form functions
// (speed_of_sound / time_step / grid_spacing) are ever called here. Sample-rate conversion is out of
// scope (DESIGN.md "Out of scope"): src_simple reports failure.
#pragma once
typedef struct {
    const float* data_in;
    float* data_out;
    long input_frames, output_frames;
    long input_frames_used, output_frames_gen;
    int end_of_input;
    double src_ratio;
} SRC_DATA;
enum { SRC_SINC_BEST_QUALITY = 0 };
inline int src_simple(SRC_DATA*, int, int) { return 1; }
inline const char* src_strerror(int) { return "libsamplerate is not part of this build"; }
