/*
 * gdext_fake_host.c -- TEST INFRASTRUCTURE: a stand-in for the Godot engine's side of the GDExtension interface, just enough to
 * load godot/libgsr_godot.so the way the engine does (dlopen, entry symbol, get_proc_address table, initialization levels), record
 * what the shim registers, and ptr-call / variant-call its methods:
 *     gdext_fake_host <libgsr_godot.so> register                                  -> prints the class + method table as JSON lines
 *     gdext_fake_host <libgsr_godot.so> render <in.bin> <out.bin>                 -> create -> resize -> upload_splats -> render -> pick -> stats
 * in.bin: int64 n, int64 w, int64 h, float vp[32], uint8 uniforms[32], float splat60[n*60];  out.bin: float rgba[w*h*4], float pick[4], gsr_stats
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../godotgaussiansplatting_b200/godot/gdextension_min.h"
#include "../include/gsr.h"

typedef struct { uint8_t *data; int64_t size; } FakePBA;                       /* the engine's PackedByteArray, as far as the shim can tell */
typedef struct { int type; union { int64_t i; double f; FakePBA a; } v; } FakeVariant;

static struct { char cls[64], parent[64]; GDExtensionClassCreationInfo2 ci; int n_methods; struct { char name[64]; int argc; int argt[8]; GDExtensionClassMethodPtrCall ptr; GDExtensionClassMethodCall call; void *ud; } m[32]; } R;
static void *g_instance;
static int g_errors;

static void f_string_name_new(GDExtensionUninitializedStringNamePtr dst, const char *s, GDExtensionBool is_static) { (void)is_static; *(const char **)dst = s; }
static void f_string_new(GDExtensionUninitializedStringPtr dst, const char *s) { *(const char **)dst = s; }
static const char *sn(GDExtensionConstStringNamePtr p) { return *(const char *const *)p; }
static void f_register_class(GDExtensionClassLibraryPtr lib, GDExtensionConstStringNamePtr name, GDExtensionConstStringNamePtr parent, const GDExtensionClassCreationInfo2 *ci) {
    (void)lib; snprintf(R.cls, sizeof R.cls, "%s", sn(name)); snprintf(R.parent, sizeof R.parent, "%s", sn(parent)); R.ci = *ci;
}
static void f_register_method(GDExtensionClassLibraryPtr lib, GDExtensionConstStringNamePtr cls, const GDExtensionClassMethodInfo *mi) {
    (void)lib; (void)cls;
    int k = R.n_methods++;
    snprintf(R.m[k].name, sizeof R.m[k].name, "%s", sn(mi->name));
    R.m[k].argc = (int)mi->argument_count; R.m[k].ptr = mi->ptrcall_func; R.m[k].call = mi->call_func; R.m[k].ud = mi->method_userdata;
    for (uint32_t i = 0; i < mi->argument_count; ++i) R.m[k].argt[i] = (int)mi->arguments_info[i].type;
}
static void f_unregister_class(GDExtensionClassLibraryPtr lib, GDExtensionConstStringNamePtr name) { (void)lib; (void)name; }
static GDExtensionObjectPtr f_construct_object(GDExtensionConstStringNamePtr cls) { (void)cls; return malloc(16); }
static void f_object_set_instance(GDExtensionObjectPtr o, GDExtensionConstStringNamePtr cls, GDExtensionClassInstancePtr inst) { (void)o; (void)cls; g_instance = inst; }
static uint8_t *f_pba_index(GDExtensionTypePtr self, GDExtensionInt i) { FakePBA *a = (FakePBA *)self; return a->size > i ? a->data + i : NULL; }
static const uint8_t *f_pba_index_const(GDExtensionConstTypePtr self, GDExtensionInt i) { const FakePBA *a = (const FakePBA *)self; return a->size > i ? a->data + i : NULL; }
static void f_print_error(const char *d, const char *fn, const char *file, int32_t line, GDExtensionBool n) { (void)file; (void)line; (void)n; ++g_errors; fprintf(stderr, "[godot print_error] %s: %s\n", fn, d); }
static void v_to_int(GDExtensionUninitializedTypePtr dst, GDExtensionVariantPtr v) { *(int64_t *)dst = ((FakeVariant *)v)->v.i; }
static void v_to_float(GDExtensionUninitializedTypePtr dst, GDExtensionVariantPtr v) { *(double *)dst = ((FakeVariant *)v)->v.f; }
static void v_to_pba(GDExtensionUninitializedTypePtr dst, GDExtensionVariantPtr v) { *(FakePBA *)dst = ((FakeVariant *)v)->v.a; }
static void v_from_int(GDExtensionUninitializedVariantPtr dst, GDExtensionTypePtr src) { ((FakeVariant *)dst)->type = 2; ((FakeVariant *)dst)->v.i = *(int64_t *)src; }
static GDExtensionTypeFromVariantConstructorFunc f_to_type(GDExtensionVariantType t) { return t == 2 ? v_to_int : t == 3 ? v_to_float : t == 29 ? v_to_pba : NULL; }
static GDExtensionVariantFromTypeConstructorFunc f_from_type(GDExtensionVariantType t) { return t == 2 ? v_from_int : NULL; }

static GDExtensionInterfaceFunctionPtr get_proc(const char *name) {
#define E(n, f) if (!strcmp(name, n)) return (GDExtensionInterfaceFunctionPtr)f
    E("string_name_new_with_latin1_chars", f_string_name_new); E("string_new_with_latin1_chars", f_string_new);
    E("classdb_register_extension_class2", f_register_class); E("classdb_register_extension_class_method", f_register_method);
    E("classdb_unregister_extension_class", f_unregister_class); E("classdb_construct_object", f_construct_object);
    E("object_set_instance", f_object_set_instance); E("packed_byte_array_operator_index", f_pba_index);
    E("packed_byte_array_operator_index_const", f_pba_index_const); E("print_error", f_print_error);
    E("get_variant_to_type_constructor", f_to_type); E("get_variant_from_type_constructor", f_from_type);
#undef E
    return NULL;
}

static int find(const char *name) { for (int k = 0; k < R.n_methods; ++k) if (!strcmp(R.m[k].name, name)) return k; fprintf(stderr, "method %s not registered\n", name); exit(3); }
static int64_t pcall(const char *name, const void *const *args) { int k = find(name); int64_t r = -1; R.m[k].ptr(R.m[k].ud, g_instance, args, &r); return r; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <libgsr_godot.so> register|render ...\n", argv[0]); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    GDExtensionInitializationFunction init = (GDExtensionInitializationFunction)dlsym(h, "gsr_gdext_init");   /* entry_symbol of addons/gsr/gsr.gdextension */
    if (!init) { fprintf(stderr, "entry symbol missing\n"); return 2; }
    GDExtensionInitialization ini;
    memset(&ini, 0, sizeof ini);
    int lib_token = 42;
    if (!init(get_proc, &lib_token, &ini)) { fprintf(stderr, "gsr_gdext_init returned false\n"); return 2; }
    for (int lvl = 0; lvl < GDEXTENSION_MAX_INITIALIZATION_LEVEL; ++lvl) if (lvl >= (int)ini.minimum_initialization_level) ini.initialize(ini.userdata, (GDExtensionInitializationLevel)lvl);
    printf("{\"class\": \"%s\", \"parent\": \"%s\", \"min_level\": %d, \"exposed\": %d, \"has_create\": %d, \"has_free\": %d}\n", R.cls, R.parent,
           (int)ini.minimum_initialization_level, (int)R.ci.is_exposed, R.ci.create_instance_func != NULL, R.ci.free_instance_func != NULL);
    for (int k = 0; k < R.n_methods; ++k) {
        printf("{\"method\": \"%s\", \"argc\": %d, \"types\": [", R.m[k].name, R.m[k].argc);
        for (int i = 0; i < R.m[k].argc; ++i) printf("%s%d", i ? ", " : "", R.m[k].argt[i]);
        printf("], \"ptrcall\": %d, \"call\": %d}\n", R.m[k].ptr != NULL, R.m[k].call != NULL);
    }
    GDExtensionObjectPtr obj = R.ci.create_instance_func(R.ci.class_userdata);   /* GsrRasterizer.new() */
    if (!obj || !g_instance) { fprintf(stderr, "create_instance failed\n"); return 2; }
    int rc_exit = 0;
    if (!strcmp(argv[2], "register")) {
        int64_t a0 = 1000, a1 = 0, a2 = 0, a3 = 10;
        const void *args[4] = {&a0, &a1, &a2, &a3};
        const long long create_rc = (long long)pcall("create", args);   /* no GPU: GSR_ERR_CUDA, reported through print_error */
        printf("{\"create_rc\": %lld, \"engine_errors\": %d}\n", create_rc, g_errors);
    } else {
        FILE *f = fopen(argv[3], "rb");
        int64_t hdr[3];
        float vp[32]; uint8_t ub[32];
        if (!f || fread(hdr, 8, 3, f) != 3 || fread(vp, 4, 32, f) != 32 || fread(ub, 1, 32, f) != 32) { fprintf(stderr, "bad input\n"); return 2; }
        const int64_t n = hdr[0], w = hdr[1], hgt = hdr[2];
        float *splats = malloc((size_t)n * 240);
        if (fread(splats, 240, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short input\n"); return 2; }
        fclose(f);
        int64_t zero = 0, ten = 10, first = 0;
        { const void *a[4] = {&n, &zero, &zero, &ten}; if (pcall("create", a)) rc_exit = 4; }
        {   /* resize through the VARIANT call path (untyped GDScript) */
            FakeVariant vw = {2, {.i = w}}, vh = {2, {.i = hgt}}, ret = {0, {.i = -1}};
            const GDExtensionConstVariantPtr va[2] = {&vw, &vh};
            GDExtensionCallError ce = {0, 0, 0};
            int k = find("resize");
            R.m[k].call(R.m[k].ud, g_instance, va, 2, &ret, &ce);
            if (ce.error != GDEXTENSION_CALL_OK || ret.type != 2 || ret.v.i != 0) rc_exit = 5;
        }
        FakePBA pb_splats = {(uint8_t *)splats, n * 240}, pb_vp = {(uint8_t *)vp, 128}, pb_ub = {ub, 32};
        { const void *a[3] = {&pb_splats, &first, &n}; if (pcall("upload_splats", a)) rc_exit = 6; }
        float *rgba = malloc((size_t)w * hgt * 16);
        FakePBA pb_out = {(uint8_t *)rgba, w * hgt * 16}, pb_none = {NULL, 0};
        double heat = 0.0;
        { const void *a[4] = {&pb_vp, &pb_ub, &heat, &pb_none}; if (pcall("render", a)) rc_exit = 7; }     /* frame stays on the device */
        { const void *a[4] = {&pb_vp, &pb_ub, &heat, &pb_out}; if (pcall("render", a)) rc_exit = 7; }      /* frame copied out */
        float pick[4] = {0, 0, 0, 0};
        FakePBA pb_pick = {(uint8_t *)pick, 16};
        int64_t tile = ((hgt / 2) / 16) * ((w + 15) / 16) + (w / 2) / 16;
        { const void *a[3] = {&tile, &heat, &pb_pick}; if (pcall("pick", a)) rc_exit = 8; }
        gsr_stats st;
        FakePBA pb_st = {(uint8_t *)&st, sizeof st};
        { const void *a[1] = {&pb_st}; if (pcall("stats", a)) rc_exit = 9; }
        int64_t fbp = pcall("framebuffer_ptr", NULL);
        printf("{\"render_rc\": %d, \"duplicates\": %llu, \"fb_ptr_nonzero\": %d, \"engine_errors\": %d}\n", rc_exit, (unsigned long long)st.duplicates, fbp != 0, g_errors);
        f = fopen(argv[4], "wb");
        fwrite(rgba, 16, (size_t)w * hgt, f); fwrite(pick, 4, 4, f); fwrite(&st, sizeof st, 1, f);
        fclose(f);
    }
    R.ci.free_instance_func(R.ci.class_userdata, g_instance);   /* RefCounted unreferenced: the context is destroyed with the instance */
    ini.deinitialize(ini.userdata, GDEXTENSION_INITIALIZATION_SCENE);
    return rc_exit;
}
