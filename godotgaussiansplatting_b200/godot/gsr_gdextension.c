/*
 * gsr_gdextension.c -- the GDExtension entry of libgsr: registers the class `GsrRasterizer` (extends RefCounted) whose methods are
 * thin ptr-calls into the C-ABI of include/gsr.h.  The GDScript class `GaussianSplattingRasterizer`
 * (util/gaussian_splatting_rasterizer.gd:2-195) keeps its public surface and forwards to an instance of this class
 * (addons/gsr/gaussian_splatting_rasterizer_gsr.gd).
 *
 * Manifest: addons/gsr/gsr.gdextension (pattern: addons/imgui-godot/imgui-godot-native.gdextension:1-13 of the reference),
 * entry_symbol = "gsr_gdext_init".
 *
 * Method table (all return the gsr status code unless noted; PackedByteArray arguments carry raw little-endian floats):
 *   create(max_splats, device, flags, dup_capacity_factor)      gsr_create          rasterizer.gd:59-114  _init / init_gpu
 *   destroy()                                                    gsr_destroy         rasterizer.gd:116-120 cleanup_gpu
 *   resize(width, height)                                        gsr_resize          rasterizer.gd:26-48   texture_size setter
 *   upload_ply_raw(vertices, nprops, first, count, time)         gsr_upload_ply_raw  ply_file.gd:28-77     loader thread
 *   upload_splats(splat60, first, count)                         gsr_upload_splats_aos  ply_file.gd:71     buffer_update
 *   render(push_constants128, uniforms32, heatmap, out_rgba)     gsr_render          rasterizer.gd:122-160 rasterize (out may be empty)
 *   pick(tile_id, heatmap, out16)                                gsr_pick            rasterizer.gd:162-171 get_splat_position
 *   stats(out)                                                   gsr_get_stats       main.gd:93-119        update_debug_info
 *   framebuffer_ptr() -> int                                     gsr_framebuffer_device_ptr  rasterizer.gd:48,101 texture_rd_rid
 */
#include <stdlib.h>
#include <string.h>

#ifdef GSR_USE_ENGINE_GDEXTENSION_HEADER
#include <gdextension_interface.h>
#else
#include "gdextension_min.h"
#endif
#include "../../include/gsr.h"

#define GSR_GDEXT_API __attribute__((visibility("default")))

static struct {
    GDExtensionClassLibraryPtr library;
    GDExtensionInterfaceStringNameNewWithLatin1Chars string_name_new;
    GDExtensionInterfaceStringNewWithLatin1Chars string_new;
    GDExtensionInterfaceClassdbRegisterExtensionClass2 register_class;
    GDExtensionInterfaceClassdbRegisterExtensionClassMethod register_method;
    GDExtensionInterfaceClassdbUnregisterExtensionClass unregister_class;
    GDExtensionInterfaceClassdbConstructObject construct_object;
    GDExtensionInterfaceObjectSetInstance object_set_instance;
    GDExtensionInterfacePackedByteArrayOperatorIndex pba_index;
    GDExtensionInterfacePackedByteArrayOperatorIndexConst pba_index_const;
    GDExtensionInterfacePrintError print_error;
    GDExtensionTypeFromVariantConstructorFunc to_int, to_float, to_pba;
    GDExtensionVariantFromTypeConstructorFunc from_int;
    uint64_t sn_class[1], sn_parent[1], sn_empty[1], str_empty[1]; /* StringName / String are pointer-sized opaque values */
} G;

typedef struct {
    gsr_ctx *ctx;
    GDExtensionObjectPtr object;
} GsrInstance;

static void report(const char *what, int rc) {
    if (rc != GSR_OK && G.print_error) G.print_error(gsr_last_error(), what, __FILE__, __LINE__, 0);
}

/* ---- argument access for ptr-calls: int = int64_t*, float = double*, PackedByteArray = the engine's opaque value ---- */
#define ARG_INT(i) (*(const int64_t *)p_args[i])
#define ARG_FLOAT(i) (*(const double *)p_args[i])
#define ARG_BYTES_CONST(i) (G.pba_index_const(p_args[i], 0))
#define ARG_BYTES(i) (G.pba_index((GDExtensionTypePtr)p_args[i], 0))
#define RET_INT(v) do { if (r_ret) *(int64_t *)r_ret = (int64_t)(v); } while (0)

typedef enum { M_CREATE, M_DESTROY, M_RESIZE, M_UPLOAD_PLY_RAW, M_UPLOAD_SPLATS, M_RENDER, M_PICK, M_STATS, M_FRAMEBUFFER_PTR, M_COUNT } MethodId;

typedef struct { const char *name; int argc; GDExtensionVariantType argt[6]; const char *argn[6]; } MethodDesc;
#define T_I GDEXTENSION_VARIANT_TYPE_INT
#define T_F GDEXTENSION_VARIANT_TYPE_FLOAT
#define T_B GDEXTENSION_VARIANT_TYPE_PACKED_BYTE_ARRAY
static const MethodDesc METHODS[M_COUNT] = {
    {"create", 4, {T_I, T_I, T_I, T_I}, {"max_splats", "device", "flags", "dup_capacity_factor"}},
    {"destroy", 0, {0}, {0}},
    {"resize", 2, {T_I, T_I}, {"width", "height"}},
    {"upload_ply_raw", 5, {T_B, T_I, T_I, T_I, T_F}, {"vertices", "nprops", "first", "count", "creation_time"}},
    {"upload_splats", 3, {T_B, T_I, T_I}, {"splat60", "first", "count"}},
    {"render", 4, {T_B, T_B, T_F, T_B}, {"push_constants", "uniforms", "heatmap_factor", "out_rgba32f"}},
    {"pick", 3, {T_I, T_F, T_B}, {"tile_id", "heatmap_factor", "out_xyzn"}},
    {"stats", 1, {T_B}, {"out_stats"}},
    {"framebuffer_ptr", 0, {0}, {0}},
};

static void method_ptrcall(void *method_userdata, GDExtensionClassInstancePtr p_instance, const GDExtensionConstTypePtr *p_args, GDExtensionTypePtr r_ret) {
    GsrInstance *self = (GsrInstance *)p_instance;
    const MethodId id = (MethodId)(intptr_t)method_userdata;
    int rc = GSR_ERR_STATE;
    if (!self) { RET_INT(GSR_ERR_INVALID); return; }
    switch (id) {
        case M_CREATE: {
            if (self->ctx) { gsr_destroy(self->ctx); self->ctx = NULL; }
            gsr_config cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.max_splats = (uint64_t)ARG_INT(0); cfg.device = (int32_t)ARG_INT(1); cfg.flags = (uint32_t)ARG_INT(2);
            cfg.dup_capacity_factor = (uint32_t)ARG_INT(3);
            rc = gsr_create(&cfg, &self->ctx);
            break;
        }
        case M_DESTROY: rc = gsr_destroy(self->ctx); self->ctx = NULL; break;
        case M_RESIZE: rc = self->ctx ? gsr_resize(self->ctx, (int32_t)ARG_INT(0), (int32_t)ARG_INT(1)) : GSR_ERR_STATE; break;
        case M_UPLOAD_PLY_RAW:
            rc = self->ctx ? gsr_upload_ply_raw(self->ctx, (const float *)ARG_BYTES_CONST(0), (uint32_t)ARG_INT(1), (uint64_t)ARG_INT(2), (uint64_t)ARG_INT(3), (float)ARG_FLOAT(4))
                           : GSR_ERR_STATE;
            break;
        case M_UPLOAD_SPLATS:
            rc = self->ctx ? gsr_upload_splats_aos(self->ctx, (const float *)ARG_BYTES_CONST(0), (uint64_t)ARG_INT(1), (uint64_t)ARG_INT(2)) : GSR_ERR_STATE;
            break;
        case M_RENDER: {
            /* out_rgba32f: a PackedByteArray of width*height*16 bytes, or an EMPTY one to keep the frame on the device
             * (the reference's Texture2DRD path; gsr_framebuffer_device_ptr / gsr_present_device hand it to the renderer) */
            float *out = (float *)ARG_BYTES(3); /* the engine returns NULL for index 0 of an empty array */
            rc = self->ctx ? gsr_render(self->ctx, (const float *)ARG_BYTES_CONST(0), ARG_BYTES_CONST(1), (float)ARG_FLOAT(2), out) : GSR_ERR_STATE;
            break;
        }
        case M_PICK: rc = self->ctx ? gsr_pick(self->ctx, (uint32_t)ARG_INT(0), (float)ARG_FLOAT(1), (float *)ARG_BYTES(2)) : GSR_ERR_STATE; break;
        case M_STATS: rc = self->ctx ? gsr_get_stats(self->ctx, (gsr_stats *)ARG_BYTES(0)) : GSR_ERR_STATE; break;
        case M_FRAMEBUFFER_PTR: RET_INT((intptr_t)(self->ctx ? gsr_framebuffer_device_ptr(self->ctx) : NULL)); return;
        default: rc = GSR_ERR_INVALID; break;
    }
    report(METHODS[id].name, rc);
    RET_INT(rc);
}

/* Variant call (untyped GDScript): unpack to the ptr-call representation, delegate, box the integer result. */
static void method_call(void *method_userdata, GDExtensionClassInstancePtr p_instance, const GDExtensionConstVariantPtr *p_args, GDExtensionInt p_argument_count,
                        GDExtensionVariantPtr r_return, GDExtensionCallError *r_error) {
    const MethodDesc *d = &METHODS[(intptr_t)method_userdata];
    if (p_argument_count != d->argc) {
        if (r_error) { r_error->error = p_argument_count < d->argc ? GDEXTENSION_CALL_ERROR_TOO_FEW_ARGUMENTS : GDEXTENSION_CALL_ERROR_TOO_MANY_ARGUMENTS; r_error->expected = d->argc; r_error->argument = 0; }
        return;
    }
    int64_t ints[6]; double floats[6]; uint64_t arrays[6][2]; /* a PackedByteArray value is two pointers wide */
    const void *typed[6];
    for (int i = 0; i < d->argc; ++i) {
        if (d->argt[i] == T_I) { G.to_int(&ints[i], (GDExtensionVariantPtr)p_args[i]); typed[i] = &ints[i]; }
        else if (d->argt[i] == T_F) { G.to_float(&floats[i], (GDExtensionVariantPtr)p_args[i]); typed[i] = &floats[i]; }
        else { G.to_pba(arrays[i], (GDExtensionVariantPtr)p_args[i]); typed[i] = arrays[i]; }
    }
    int64_t ret = 0;
    method_ptrcall(method_userdata, p_instance, typed, &ret);
    if (r_return) G.from_int(r_return, &ret);
    if (r_error) r_error->error = GDEXTENSION_CALL_OK;
}

static GDExtensionObjectPtr class_create_instance(void *class_userdata) {
    (void)class_userdata;
    GsrInstance *self = (GsrInstance *)calloc(1, sizeof *self);
    if (!self) return NULL;
    self->object = G.construct_object(G.sn_parent);
    G.object_set_instance(self->object, G.sn_class, self);
    return self->object;
}

static void class_free_instance(void *class_userdata, GDExtensionClassInstancePtr p_instance) {
    (void)class_userdata;
    GsrInstance *self = (GsrInstance *)p_instance;
    if (!self) return;
    if (self->ctx) gsr_destroy(self->ctx);   /* NOTIFICATION_PREDELETE -> deletion queue of render_context.gd:40-44 */
    free(self);
}

static void register_everything(void) {
    GDExtensionClassCreationInfo2 ci;
    memset(&ci, 0, sizeof ci);
    ci.is_exposed = 1;
    ci.create_instance_func = class_create_instance;
    ci.free_instance_func = class_free_instance;
    G.register_class(G.library, G.sn_class, G.sn_parent, &ci);
    for (int m = 0; m < M_COUNT; ++m) {
        const MethodDesc *d = &METHODS[m];
        uint64_t name[1], argn[6][1];
        GDExtensionPropertyInfo args[6], ret;
        GDExtensionClassMethodArgumentMetadata meta[6];
        G.string_name_new(name, d->name, 1);
        for (int i = 0; i < d->argc; ++i) {
            G.string_name_new(argn[i], d->argn[i], 1);
            args[i].type = d->argt[i]; args[i].name = argn[i]; args[i].class_name = G.sn_empty; args[i].hint = 0; args[i].hint_string = G.str_empty;
            args[i].usage = 6; /* PROPERTY_USAGE_DEFAULT */
            meta[i] = d->argt[i] == T_I ? GDEXTENSION_METHOD_ARGUMENT_METADATA_INT_IS_INT64 : (d->argt[i] == T_F ? GDEXTENSION_METHOD_ARGUMENT_METADATA_REAL_IS_DOUBLE : GDEXTENSION_METHOD_ARGUMENT_METADATA_NONE);
        }
        ret.type = GDEXTENSION_VARIANT_TYPE_INT; ret.name = G.sn_empty; ret.class_name = G.sn_empty; ret.hint = 0; ret.hint_string = G.str_empty; ret.usage = 6;
        GDExtensionClassMethodInfo mi;
        memset(&mi, 0, sizeof mi);
        mi.name = name; mi.method_userdata = (void *)(intptr_t)m; mi.call_func = method_call; mi.ptrcall_func = method_ptrcall;
        mi.method_flags = GDEXTENSION_METHOD_FLAGS_DEFAULT; mi.has_return_value = 1; mi.return_value_info = &ret;
        mi.return_value_metadata = GDEXTENSION_METHOD_ARGUMENT_METADATA_INT_IS_INT64;
        mi.argument_count = (uint32_t)d->argc; mi.arguments_info = args; mi.arguments_metadata = meta;
        G.register_method(G.library, G.sn_class, &mi);
    }
}

static void on_initialize(void *userdata, GDExtensionInitializationLevel level) {
    (void)userdata;
    if (level == GDEXTENSION_INITIALIZATION_SCENE) register_everything();
}
static void on_deinitialize(void *userdata, GDExtensionInitializationLevel level) {
    (void)userdata;
    if (level == GDEXTENSION_INITIALIZATION_SCENE && G.unregister_class) G.unregister_class(G.library, G.sn_class);
}

GSR_GDEXT_API GDExtensionBool gsr_gdext_init(GDExtensionInterfaceGetProcAddress get_proc, GDExtensionClassLibraryPtr library, GDExtensionInitialization *r_init) {
    if (!get_proc || !r_init) return 0;
    memset(&G, 0, sizeof G);
    G.library = library;
#define LOAD(field, type, name) do { G.field = (type)get_proc(name); if (!G.field) return 0; } while (0)
    LOAD(string_name_new, GDExtensionInterfaceStringNameNewWithLatin1Chars, "string_name_new_with_latin1_chars");
    LOAD(string_new, GDExtensionInterfaceStringNewWithLatin1Chars, "string_new_with_latin1_chars");
    LOAD(register_class, GDExtensionInterfaceClassdbRegisterExtensionClass2, "classdb_register_extension_class2");
    LOAD(register_method, GDExtensionInterfaceClassdbRegisterExtensionClassMethod, "classdb_register_extension_class_method");
    LOAD(construct_object, GDExtensionInterfaceClassdbConstructObject, "classdb_construct_object");
    LOAD(object_set_instance, GDExtensionInterfaceObjectSetInstance, "object_set_instance");
    LOAD(pba_index, GDExtensionInterfacePackedByteArrayOperatorIndex, "packed_byte_array_operator_index");
    LOAD(pba_index_const, GDExtensionInterfacePackedByteArrayOperatorIndexConst, "packed_byte_array_operator_index_const");
#undef LOAD
    G.unregister_class = (GDExtensionInterfaceClassdbUnregisterExtensionClass)get_proc("classdb_unregister_extension_class");
    G.print_error = (GDExtensionInterfacePrintError)get_proc("print_error");
    GDExtensionInterfaceGetVariantToTypeConstructor to_type = (GDExtensionInterfaceGetVariantToTypeConstructor)get_proc("get_variant_to_type_constructor");
    GDExtensionInterfaceGetVariantFromTypeConstructor from_type = (GDExtensionInterfaceGetVariantFromTypeConstructor)get_proc("get_variant_from_type_constructor");
    if (!to_type || !from_type) return 0;
    G.to_int = to_type(GDEXTENSION_VARIANT_TYPE_INT); G.to_float = to_type(GDEXTENSION_VARIANT_TYPE_FLOAT); G.to_pba = to_type(GDEXTENSION_VARIANT_TYPE_PACKED_BYTE_ARRAY);
    G.from_int = from_type(GDEXTENSION_VARIANT_TYPE_INT);
    if (!G.to_int || !G.to_float || !G.to_pba || !G.from_int) return 0;
    G.string_name_new(G.sn_class, "GsrRasterizer", 1);
    G.string_name_new(G.sn_parent, "RefCounted", 1);
    G.string_name_new(G.sn_empty, "", 1);
    G.string_new(G.str_empty, "");
    r_init->minimum_initialization_level = GDEXTENSION_INITIALIZATION_SCENE;
    r_init->userdata = NULL;
    r_init->initialize = on_initialize;
    r_init->deinitialize = on_deinitialize;
    return 1;
}
