/* glava_b200: smoothing / transform parameters (GLava's shipped values) */
#define ROUND_FORMULA sinusoidal
#define SAMPLE_MODE average
#define SAMPLE_HYBRID_WEIGHT 0.65
#define SAMPLE_SCALE 8
#define SAMPLE_RANGE 0.9
#request setfftscale 10.2
#request setfftcutoff 0.3
#request setavgframes 5
#request setavgwindow true
#request setgravitystep 4.2
#request setsmoothfactor 0.025
#request setsmoothpass true
