/* TEST INFRASTRUCTURE — not product code.
 * render.c references four X11 helpers (xwin.h:12,13,14,18) whose implementation
 * (xwin.c) needs Xlib, which this image does not have.  None is reachable from the
 * CPU transforms the shim exports; these no-op definitions only satisfy the linker. */
#include <stdbool.h>
struct gl_wcb;
void xwin_wait_for_wm(void) {}
void xwin_assign_icon_bmp(struct gl_wcb* wcb, void* impl, const char* path) { (void) wcb; (void) impl; (void) path; }
bool xwin_should_render(struct gl_wcb* wcb, void* impl) { (void) wcb; (void) impl; return true; }
unsigned int xwin_copyglbg(void* rd, unsigned int texture) { (void) rd; return texture; }
