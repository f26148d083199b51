/* TEST INFRASTRUCTURE — not product code.
 *
 * A display-less stand-in for libX11.so.6 / libXext.so.6, just large enough for the Mesa software GL that ships inside
 * Nsight Compute in this image (…/nsight-compute/…/Mesa/libGL.so.1.5.0: Mesa 18.1.9, gallium "libgl-xlib" target with
 * llvmpipe + softpipe) to create an OpenGL 3.3 core context on a pbuffer WITHOUT an X server.  GLava names llvmpipe as
 * its software floor (reference README.md:121); with this the reference's own shaders are compiled by a real GLSL
 * compiler and rasterised by llvmpipe, which is what pins the raster half of the oracle.
 *
 * Only what Mesa's xlib winsys / GLX emulation touches is provided: one screen, one 24-bit TrueColor visual, no
 * MIT-SHM, drawing requests are no-ops (frames are read back with glReadPixels from FBOs, nothing is ever presented).
 * The structure layouts are Xlib's stable public ABI (Xlib.h / Xutil.h / Xlibint.h), restated here because the image
 * has no X11 headers.  Built by oracle/Makefile into oracle/_ref/fakex/.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID;
typedef XID Window, Drawable, Pixmap, Colormap, Font, VisualID;
typedef int Bool, Status;
typedef char* XPointer;
typedef struct _XGC* GC;

typedef struct _XExtData { int number; struct _XExtData* next; int (*free_private)(struct _XExtData*); XPointer private_data; } XExtData;
typedef struct { int extension, major_opcode, first_event, first_error; } XExtCodes;
typedef struct { XExtData* ext_data; VisualID visualid; int class_; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb, map_entries; } Visual;
typedef struct { int depth, nvisuals; Visual* visuals; } Depth;
typedef struct { XExtData* ext_data; int depth, bits_per_pixel, scanline_pad; } ScreenFormat;
struct _XDisplay;
typedef struct {
    XExtData* ext_data; struct _XDisplay* display; Window root; int width, height, mwidth, mheight; int ndepths; Depth* depths;
    int root_depth; Visual* root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel, black_pixel;
    int max_maps, min_maps, backing_store; Bool save_unders; long root_input_mask;
} Screen;

/* Xlibint.h `_XExten`: Mesa's GLX emulation registers a close-display hook by reaching into dpy->ext_procs */
typedef struct _XExten {
    struct _XExten* next; XExtCodes codes;
    void *create_GC, *copy_GC, *flush_GC, *free_GC, *create_Font, *free_Font, *close_display, *error, *error_string;
    char* name; void *error_values, *before_flush; struct _XExten* next_flush;
} _XExtension;

/* Xlibint.h `struct _XDisplay` up to ext_procs (the public `_XPrivDisplay` of Xlib.h is its prefix), then slack */
typedef struct _XDisplay {
    XExtData* ext_data; void* free_funcs; int fd; int conn_checker; int proto_major_version, proto_minor_version; char* vendor;
    XID resource_base, resource_mask, resource_id; int resource_shift; XID (*resource_alloc)(struct _XDisplay*);
    int byte_order, bitmap_unit, bitmap_pad, bitmap_bit_order; int nformats; ScreenFormat* pixmap_format; int vnumber, release;
    void *head, *tail; int qlen; unsigned long last_request_read, request; char *last_req, *buffer, *bufptr, *bufmax;
    unsigned max_request_size; void* db; int (*synchandler)(struct _XDisplay*); char* display_name; int default_screen, nscreens;
    Screen* screens; unsigned long motion_buffer; volatile unsigned long flags; int min_keycode, max_keycode; void* keysyms;
    void* modifiermap; int keysyms_per_keycode; char* xdefaults; char* scratch_buffer; unsigned long scratch_length;
    int ext_number; _XExtension* ext_procs;
    char slack[16384];                                   /* everything after that reads as zero / NULL */
} Display;

typedef struct { Visual* visual; VisualID visualid; int screen, depth, class_; unsigned long red_mask, green_mask, blue_mask; int colormap_size, bits_per_rgb; } XVisualInfo;
typedef struct {
    int x, y, width, height, border_width, depth; Visual* visual; Window root; int class_, bit_gravity, win_gravity, backing_store;
    unsigned long backing_planes, backing_pixel; Bool save_under; Colormap colormap; Bool map_installed; int map_state;
    long all_event_masks, your_event_mask, do_not_propagate_mask; Bool override_redirect; Screen* screen;
} XWindowAttributes;
typedef struct _XImage {
    int width, height, xoffset, format; char* data; int byte_order, bitmap_unit, bitmap_bit_order, bitmap_pad, depth, bytes_per_line, bits_per_pixel;
    unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
    struct funcs {
        struct _XImage* (*create_image)(void); int (*destroy_image)(struct _XImage*); unsigned long (*get_pixel)(struct _XImage*, int, int);
        int (*put_pixel)(struct _XImage*, int, int, unsigned long); struct _XImage* (*sub_image)(struct _XImage*, int, int, unsigned, unsigned);
        int (*add_pixel)(struct _XImage*, long);
    } f;
} XImage;

#define FX_W 4096
#define FX_H 4096
static Visual  fx_visual = { NULL, 0x21, 4 /* TrueColor */, 0xff0000, 0x00ff00, 0x0000ff, 8, 256 };
static Depth   fx_depth  = { 24, 1, &fx_visual };
static ScreenFormat fx_format = { NULL, 24, 32, 32 };
static Screen  fx_screen;
static Display fx_display;
static struct _XGC { int dummy; } fx_gc;

/* Not Xlib: the handle the GL harness passes to glX* (there is no server to connect to, so no XOpenDisplay either). */
Display* fakex_display(void) {
    if (!fx_display.screens) {
        memset(&fx_display, 0, sizeof(fx_display));
        fx_screen = (Screen) { .display = &fx_display, .root = 0x100, .width = FX_W, .height = FX_H, .mwidth = 1000, .mheight = 1000,
                               .ndepths = 1, .depths = &fx_depth, .root_depth = 24, .root_visual = &fx_visual,
                               .default_gc = &fx_gc, .cmap = 0x20, .white_pixel = 0xffffff, .black_pixel = 0, .max_maps = 1, .min_maps = 1 };
        fx_display.fd = -1; fx_display.proto_major_version = 11; fx_display.vendor = (char*) "glava_b200 fake X (no server)";
        fx_display.byte_order = 0 /* LSBFirst */; fx_display.bitmap_unit = 32; fx_display.bitmap_pad = 32; fx_display.bitmap_bit_order = 0;
        fx_display.nformats = 1; fx_display.pixmap_format = &fx_format; fx_display.vnumber = 11; fx_display.release = 1;
        fx_display.display_name = (char*) ":fake"; fx_display.default_screen = 0; fx_display.nscreens = 1; fx_display.screens = &fx_screen;
        fx_display.max_request_size = 65535;
    }
    return &fx_display;
}

void (*_XLockMutex_fn)(void*) = NULL;
void (*_XUnlockMutex_fn)(void*) = NULL;
void* _Xglobal_lock = NULL;

static XID fx_next_id = 0x400000;
static int fx_trace = -1;
#define TRACE(name) do { if (fx_trace < 0) fx_trace = getenv("FAKEX_TRACE") ? 1 : 0; if (fx_trace) fprintf(stderr, "[fakex] %s\n", name); } while (0)

XExtCodes* XAddExtension(Display* dpy) {
    TRACE("XAddExtension");
    _XExtension* e = calloc(1, sizeof(*e));
    e->codes.extension = dpy->ext_number++;
    e->next = dpy->ext_procs; dpy->ext_procs = e;
    return &e->codes;
}
Bool XQueryExtension(Display* d, const char* name, int* op, int* ev, int* er) { TRACE("XQueryExtension"); (void) d; (void) name; if (op) *op = 0; if (ev) *ev = 0; if (er) *er = 0; return 0; }
Colormap XCreateColormap(Display* d, Window w, Visual* v, int alloc) { TRACE("XCreateColormap"); (void) d; (void) w; (void) v; (void) alloc; return fx_next_id++; }
GC XCreateGC(Display* d, Drawable dr, unsigned long mask, void* values) { TRACE("XCreateGC"); (void) d; (void) dr; (void) mask; (void) values; return calloc(1, 128); }
int XFreeGC(Display* d, GC gc) { (void) d; free(gc); return 1; }
static int fx_destroy_image(XImage* img) { if (img) { free(img->data); free(img); } return 1; }
XImage* XCreateImage(Display* d, Visual* v, unsigned depth, int format, int offset, char* data, unsigned w, unsigned h, int pad, int bpl) {
    TRACE("XCreateImage"); (void) d;
    XImage* img = calloc(1, sizeof(*img));
    img->width = (int) w; img->height = (int) h; img->xoffset = offset; img->format = format; img->data = data;
    img->byte_order = 0; img->bitmap_unit = 32; img->bitmap_bit_order = 0; img->bitmap_pad = pad; img->depth = (int) depth;
    img->bits_per_pixel = depth > 16 ? 32 : (depth > 8 ? 16 : (depth > 1 ? 8 : 1));
    img->bytes_per_line = bpl ? bpl : (int) (((size_t) w * img->bits_per_pixel + (pad ? pad : 8) - 1) / (pad ? pad : 8)) * ((pad ? pad : 8) / 8);
    if (v) { img->red_mask = v->red_mask; img->green_mask = v->green_mask; img->blue_mask = v->blue_mask; }
    img->f.destroy_image = fx_destroy_image;
    return img;
}
Pixmap XCreatePixmap(Display* d, Drawable dr, unsigned w, unsigned h, unsigned depth) { TRACE("XCreatePixmap"); (void) d; (void) dr; (void) w; (void) h; (void) depth; return fx_next_id++; }
int XFreePixmap(Display* d, Pixmap p) { (void) d; (void) p; return 1; }
int XDrawString16(Display* d, Drawable dr, GC gc, int x, int y, const void* s, int n) { (void) d; (void) dr; (void) gc; (void) x; (void) y; (void) s; (void) n; return 0; }
int XFillRectangle(Display* d, Drawable dr, GC gc, int x, int y, unsigned w, unsigned h) { (void) d; (void) dr; (void) gc; (void) x; (void) y; (void) w; (void) h; return 0; }
int XFlush(Display* d) { (void) d; return 1; }
int XSync(Display* d, Bool discard) { (void) d; (void) discard; return 1; }
int XFree(void* p) { free(p); return 1; }
void* XQueryFont(Display* d, XID id) { (void) d; (void) id; return NULL; }
int XFreeFontInfo(char** names, void* info, int n) { (void) names; (void) info; (void) n; return 1; }
Status XGetGeometry(Display* d, Drawable dr, Window* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* bw, unsigned* depth) {
    TRACE("XGetGeometry"); (void) d; (void) dr;
    if (root) *root = fx_screen.root; if (x) *x = 0; if (y) *y = 0; if (w) *w = 64; if (h) *h = 64; if (bw) *bw = 0; if (depth) *depth = 24;
    return 1;
}
XImage* XGetImage(Display* d, Drawable dr, int x, int y, unsigned w, unsigned h, unsigned long mask, int format) {
    (void) dr; (void) x; (void) y; (void) mask;
    XImage* img = XCreateImage(d, &fx_visual, 24, format, 0, NULL, w, h, 32, 0);
    img->data = calloc((size_t) img->bytes_per_line, h ? h : 1);
    return img;
}
XVisualInfo* XGetVisualInfo(Display* d, long mask, XVisualInfo* t, int* n) {
    TRACE("XGetVisualInfo"); (void) d;
    const XVisualInfo mine = { &fx_visual, fx_visual.visualid, 0, 24, 4, fx_visual.red_mask, fx_visual.green_mask, fx_visual.blue_mask, 256, 8 };
    int ok = 1;
    if (t) {
        if ((mask & 0x1) && t->visualid != mine.visualid) ok = 0;      /* VisualIDMask */
        if ((mask & 0x2) && t->screen != 0) ok = 0;                    /* VisualScreenMask */
        if ((mask & 0x4) && t->depth != mine.depth) ok = 0;            /* VisualDepthMask */
        if ((mask & 0x8) && t->class_ != mine.class_) ok = 0;          /* VisualClassMask */
    }
    if (!ok) { *n = 0; return NULL; }
    XVisualInfo* out = malloc(sizeof(*out)); *out = mine; *n = 1;
    return out;
}
Status XGetWindowAttributes(Display* d, Window w, XWindowAttributes* a) {
    TRACE("XGetWindowAttributes"); (void) d; (void) w;
    memset(a, 0, sizeof(*a));
    a->width = 64; a->height = 64; a->depth = 24; a->visual = &fx_visual; a->root = fx_screen.root; a->class_ = 1 /* InputOutput */;
    a->colormap = fx_screen.cmap; a->map_installed = 1; a->map_state = 2 /* IsViewable */; a->screen = &fx_screen;
    return 1;
}
int XPutImage(Display* d, Drawable dr, GC gc, XImage* img, int sx, int sy, int dx, int dy, unsigned w, unsigned h) { (void) d; (void) dr; (void) gc; (void) img; (void) sx; (void) sy; (void) dx; (void) dy; (void) w; (void) h; return 0; }
void* XSetErrorHandler(void* h) { (void) h; return NULL; }
int XSetForeground(Display* d, GC gc, unsigned long px) { (void) d; (void) gc; (void) px; return 1; }
int XSetFunction(Display* d, GC gc, int fn) { (void) d; (void) gc; (void) fn; return 1; }
void* XSynchronize(Display* d, Bool on) { (void) d; (void) on; return NULL; }
/* MIT-SHM (libXext): reported absent through XQueryExtension; the entry points exist for the link only */
Bool XShmAttach(Display* d, void* info) { (void) d; (void) info; return 0; }
XImage* XShmCreateImage(Display* d, Visual* v, unsigned depth, int format, char* data, void* info, unsigned w, unsigned h) { (void) d; (void) v; (void) depth; (void) format; (void) data; (void) info; (void) w; (void) h; return NULL; }
Bool XShmPutImage(Display* d, Drawable dr, GC gc, XImage* img, int sx, int sy, int dx, int dy, unsigned w, unsigned h, Bool ev) { (void) d; (void) dr; (void) gc; (void) img; (void) sx; (void) sy; (void) dx; (void) dy; (void) w; (void) h; (void) ev; return 0; }
