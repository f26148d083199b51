/* glava_b200 entry file: same `#request` surface as GLava's rc.glsl, restricted to the requests
   that configure the PCM -> spectrum -> pixels path.  Values are GLava's shipped defaults. */

/* visualiser module: bars | radial | circle | graph | wave */
#request mod bars
/* mono input mirrored to both channels */
#request setmirror false
/* "native" premultiplies alpha in the final stage; "none" / "xroot" do not */
#request setopacity "native"
/* framebuffer geometry: x y width height (x, y unused here) */
#request setgeometry 0 0 800 600
/* PCM ring length per channel (floats); the FFT is (bufsize / 2)-point complex */
#request setbufsize 4096
/* FIFO chunk size: samplesize / 4 new stereo frames per update */
#request setsamplesize 1024
#request setsamplerate 22050
/* true: gravity / average / smoothing as 16-bit texture passes; false: float chain of render.c */
#request setaccelfft true
#request setinterpolate false
#request setbufscale 1
