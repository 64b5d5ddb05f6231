@tool
class_name GaussianSplattingRasterizer extends Resource
## Drop-in body for util/gaussian_splatting_rasterizer.gd of 2Retr0/GodotGaussianSplatting: same class name, same members, same call
## sites (main.gd:121-156), but the six compute-shader dispatches of rasterize() (rasterizer.gd:122-160) run in libgsr on a B200
## through the GDExtension class GsrRasterizer (godotgaussiansplatting_b200/godot/gsr_gdextension.c).  Not executed in this repository
## (no Godot in the build image); it documents the binding a maintainer adds.  Every member cites the reference line it replaces.

const TILE_SIZE := 16                                    # rasterizer.gd:4
signal loaded                                            # rasterizer.gd:10

var native := GsrRasterizer.new()
var created := false
var frame_bytes := PackedByteArray()                     # RGBA32F frame for texture_update (rasterizer.gd:92)
var render_texture : Texture2DRD
var texture_rid : RID
var point_cloud : PlyFile
var camera : Camera3D
var camera_projection : Projection
var camera_transform : Projection
var camera_push_constants : PackedByteArray
var tile_dims := Vector2i.ZERO
var texture_size : Vector2i :                            # rasterizer.gd:26-48
	set(value):
		texture_size = (value * render_scale[0]).max(Vector2i.ONE)
		tile_dims = (texture_size + Vector2i.ONE*(TILE_SIZE - 1)) / TILE_SIZE
		if not created: return
		_check(native.resize(texture_size.x, texture_size.y), 'gsr_resize')
		frame_bytes.resize(texture_size.x * texture_size.y * 16)
		_make_texture()
var load_thread := Thread.new()
var is_loaded := false
var should_enable_heatmap := [false]
var render_scale := [1.0]
var model_scale := [1.0]
var should_terminate_thread : Array[bool] = [false]
var num_splats_loaded : Array[int] = [0]
var basis_override := Basis.IDENTITY

func _init(point_cloud : PlyFile, output_texture_size : Vector2i, render_texture : Texture2DRD, camera : Camera3D) -> void:   # :59
	self.point_cloud = point_cloud
	self.texture_size = output_texture_size
	self.render_texture = render_texture
	self.camera = camera

func _check(rc : int, what : String) -> void:
	if rc != 0: push_error('%s failed with gsr status %d' % [what, rc])

func _make_texture() -> void:                            # rasterizer.gd:41,48,101: the Texture2DRD main.gdshader samples
	var rd := RenderingServer.get_rendering_device()
	var fmt := RDTextureFormat.new()
	fmt.format = RenderingDevice.DATA_FORMAT_R32G32B32A32_SFLOAT
	fmt.width = texture_size.x; fmt.height = texture_size.y
	fmt.usage_bits = RenderingDevice.TEXTURE_USAGE_SAMPLING_BIT | RenderingDevice.TEXTURE_USAGE_CAN_UPDATE_BIT
	texture_rid = rd.texture_create(fmt, RDTextureView.new(), [])
	render_texture.texture_rd_rid = texture_rid

func init_gpu() -> void:                                 # rasterizer.gd:65-114
	assert(render_texture, 'An output Texture2DRD must be set!')
	_check(native.create(point_cloud.size, 0, 0, 10), 'gsr_create')   # device 0, default flags, capacity factor 10 (:79; grows on demand)
	created = true
	self.texture_size = texture_size
	should_terminate_thread[0] = false
	num_splats_loaded[0] = 0
	load_thread.start(_load.bind())

func _load() -> void:
	# PlyFile.load_gaussian_splats (ply_file.gd:28-77) without its per-splat GDScript loop: the raw vertex bytes go to the device,
	# which runs the exp / sigmoid / quaternion -> covariance / SH interleave of :44-69 itself (gsr_upload_ply_raw)
	var nprops := point_cloud.properties.size()
	var stride := maxi(1, point_cloud.size / 1000)      # ply_file.gd:36
	var i := 0
	while i * stride < point_cloud.size and not should_terminate_thread[0]:
		var first := i * stride
		var count := mini(stride, point_cloud.size - first)
		var bytes := point_cloud.vertices.slice(first * nprops * 4, (first + count) * nprops * 4)
		_check(native.upload_ply_raw(bytes, nprops, first, count, Time.get_ticks_msec() * 1e-3), 'gsr_upload_ply_raw')
		num_splats_loaded[0] = first + count
		i += 1
	is_loaded = true
	loaded.emit.call_deferred()

func cleanup_gpu() -> void:                              # rasterizer.gd:116-120
	should_terminate_thread[0] = true
	if load_thread.is_started(): load_thread.wait_to_finish()
	native.destroy()
	created = false
	if render_texture: render_texture.texture_rd_rid = RID()

func rasterize() -> void:                                # rasterizer.gd:122-160
	if not created: init_gpu()
	var camera_pos := basis_override * camera.global_position
	var uniforms := RenderingContext.create_push_constant([-camera_pos.x, -camera_pos.y, camera_pos.z, model_scale[0], texture_size.x, texture_size.y, Time.get_ticks_msec()*1e-3])   # :126
	_check(native.render(camera_push_constants, uniforms, float(should_enable_heatmap[0]), frame_bytes), 'gsr_render')
	RenderingServer.get_rendering_device().texture_update(texture_rid, 0, frame_bytes)

func get_splat_position(screen_position : Vector2i) -> Vector3:   # rasterizer.gd:162-171
	var tile : Vector2i = Vector2i(screen_position * render_scale[0]) / TILE_SIZE
	var out := PackedByteArray(); out.resize(16)
	_check(native.pick(tile.y*tile_dims.x + tile.x, float(should_enable_heatmap[0]), out), 'gsr_pick')
	var r := out.to_float32_array()
	if r[3] == 0: return Vector3.INF
	return basis_override.inverse() * Vector3(-r[0], -r[1], r[2])

func update_camera_matrices() -> bool:                   # rasterizer.gd:175-195 (unchanged arithmetic)
	var proj := camera.get_camera_projection()
	var view := Projection(Transform3D(basis_override * camera.get_camera_transform().basis, basis_override * camera.get_camera_transform().origin))
	if view == camera_transform and proj == camera_projection: return false
	camera_projection = proj
	camera_transform = view
	camera_push_constants = RenderingContext.create_push_constant([
		-view.x[0], view.y[0], -view.z[0], 0.0, -view.x[1], view.y[1], -view.z[1], 0.0,
		 view.x[2], -view.y[2], view.z[2], 0.0, -view.w.dot(view.x), view.w.dot(view.y), -view.w.dot(view.z), 1.0,
		 proj.x[0], proj.x[1], proj.x[2], 0.0, proj.y[0], proj.y[1], proj.y[2], 0.0,
		 proj.z[0], proj.z[1], proj.z[2], -1.0, proj.w[0], proj.w[1], proj.w[2], 0.0])
	return true

func stats() -> Dictionary:                              # main.gd:93-119 update_debug_info: M, overflow, stage times
	var out := PackedByteArray(); out.resize(104)   # sizeof(gsr_stats)
	_check(native.stats(out), 'gsr_get_stats')
	return {'duplicates': out.decode_u64(8), 'overflow': out.decode_u32(40) != 0,
			'stage_ms': [out.decode_float(72), out.decode_float(76), out.decode_float(80), out.decode_float(84), out.decode_float(88)]}
