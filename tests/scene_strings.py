"""Small inline .prc scenes for unit-level tests (the analogue of the reference's src/tests/testscene.inl)."""

# every material plugin on the scoped path; rough variants with roughness 0.164 as in reference src/tests/materials.cpp:37-47
MATERIAL_ZOO = """
(scene :name 'zoo' :render_width 32 :render_height 32 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 8)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,4, 0,0,0,1])
 (emission :name 'em' :type 'standard' :radiance (illuminant "D65"))
 (material :name 'm_diffuse' :type 'diffuse' :albedo (refl 0.8 0.2 0.3))
 (material :name 'm_glass' :type 'glass' :index 1.55)
 (material :name 'm_glass_bk7' :type 'glass' :index (lookup_index "bk7"))
 (material :name 'm_conductor' :type 'conductor' :eta 0.051585 :k 3.9046)
 (material :name 'm_roughconductor' :type 'roughconductor' :eta 0.051585 :k 3.9046 :roughness 0.164)
 (material :name 'm_roughconductor_aniso' :type 'roughconductor' :eta 0.2 :k 3.0 :roughness_x 0.1 :roughness_y 0.3)
 (material :name 'm_roughglass' :type 'roughglass' :index 1.55 :roughness 0.164)
 (material :name 'm_roughglass_bk7' :type 'glass' :index (lookup_index "bk7") :roughness 0.05)
 (material :name 'm_principled' :type 'principled' :base (refl 0.6 0.5 0.3) :roughness 0.164 :metallic 0.3 :sheen 0.2 :clearcoat 0.4)
 (material :name 'm_principled_trans' :type 'principled' :base (refl 0.9 0.9 0.9) :roughness 0.3 :specular_transmission 0.8 :diffuse_transmission 0.2)
 (entity :name 'floor' :type 'plane' :centering true :width 4 :height 4 :material 'm_diffuse' :position [0,0,-1])
 (entity :name 's0' :type 'sphere' :radius 0.4 :material 'm_glass' :position [-1.2,0.8,0])
 (entity :name 's1' :type 'sphere' :radius 0.4 :material 'm_glass_bk7' :position [-0.4,0.8,0])
 (entity :name 's2' :type 'sphere' :radius 0.4 :material 'm_conductor' :position [0.4,0.8,0])
 (entity :name 's3' :type 'sphere' :radius 0.4 :material 'm_roughconductor' :position [1.2,0.8,0])
 (entity :name 's4' :type 'sphere' :radius 0.4 :material 'm_roughconductor_aniso' :position [-1.2,-0.2,0])
 (entity :name 's5' :type 'sphere' :radius 0.4 :material 'm_roughglass' :position [-0.4,-0.2,0])
 (entity :name 's6' :type 'sphere' :radius 0.4 :material 'm_roughglass_bk7' :position [0.4,-0.2,0])
 (entity :name 's7' :type 'sphere' :radius 0.4 :material 'm_principled' :position [1.2,-0.2,0])
 (entity :name 's8' :type 'sphere' :radius 0.4 :material 'm_principled_trans' :position [0,-1.2,0])
 (entity :name 'lamp' :type 'plane' :centering true :width 1 :height 1 :material 'm_diffuse' :emission 'em' :transform [1,0,0,1.2, 0,-1,0,1.2, 0,0,-1,3, 0,0,0,1])
 (light :type 'env' :radiance (illuminant "D65"))
)
"""

# white furnace (reference src/tests/python/whitefurnance.py:142-195): unit sphere, albedo 1, constant env radiance 1
FURNACE = """
(scene :name 'furnace' :render_width 48 :render_height 48 :camera 'Camera' :spectral_hero %(hero)s
 (integrator :type 'direct' :max_ray_depth 4)
 (sampler :slot 'aa' :type 'random' :sample_count 8)
 (filter :slot 'pixel' :type 'block' :radius 0)
 (spectral_mapper :type 'random')
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.01 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,3, 0,0,0,1])
 (material :name 'white' :type 'diffuse' :albedo 1)
 (entity :name 'ball' :type 'sphere' :radius 1 :material 'white')
 (light :type 'env' :radiance 1)
)
"""

# (the lamp of the zoo scenes sits beside the camera's view: in front of it, as in round 1, it hid the spheres from the film tests)

# sky / sun infinite lights: the Hosek-Wilkie sky WITHOUT ground extension and with a rotated frame, a cone
# sun given by direction, a delta sun (radius 0, hasDeltaDistribution) given by date/time, over diffuse / glossy / glass spheres;
# together with scenes/c4c_complex.prc (extended sky + cone sun by hour) this covers every branch of sky.cpp / sun.cpp
SKYSUN_ZOO = """
(scene :name 'skysun' :render_width 48 :render_height 48 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 6)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (camera :name 'Camera' :type 'standard' :width 1.4 :height 1.4 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,0.7071068,0.7071068,3, 0,-0.7071068,0.7071068,3, 0,0,0,1])
 (material :name 'm_ground' :type 'diffuse' :albedo (refl 0.5 0.45 0.4))
 (material :name 'm_diffuse' :type 'diffuse' :albedo (refl 0.8 0.2 0.3))
 (material :name 'm_metal' :type 'roughconductor' :eta 0.2 :k 3.0 :roughness 0.2)
 (material :name 'm_glass' :type 'glass' :index (lookup_index "bk7"))
 (entity :name 'ground' :type 'plane' :centering true :width 8 :height 8 :material 'm_ground' :position [0,0,0])
 (entity :name 's0' :type 'sphere' :radius 0.5 :material 'm_diffuse' :position [-1.1,0,0.5])
 (entity :name 's1' :type 'sphere' :radius 0.5 :material 'm_metal' :position [0,0,0.5])
 (entity :name 's2' :type 'sphere' :radius 0.5 :material 'm_glass' :position [1.1,0,0.5])
 (light :name 'sky' :type 'sky' :turbidity 4.5 :albedo 0.3 :extend false :elevation 0.6 :azimuth 2.0
   :azimuth_resolution 64 :elevation_resolution 32 :transform [0.9800666,-0.1986693,0,0, 0.1986693,0.9800666,0,0, 0,0,1,0, 0,0,0,1])
 (light :name 'sun' :type 'sun' :turbidity 4.5 :radius 8 :direction [0.4,-0.3,0.85])
 (light :name 'sun_delta' :type 'sun' :turbidity 2 :radius 0 :power_scale 0.5 :year 2021 :month 7 :day 14 :hour 10 :minute 30
   :latitude 48.2 :longitude 16.4 :timezone 2)
)
"""

# the cheap materials of the other example scenes -- delta mirror (mirror.cpp) and improved Oren-Nayar
# (orennayar.cpp; roughness 0 must reduce to its albedo / pi like Lambert without the two-sided handling)
MATERIAL_ZOO2 = """
(scene :name 'zoo2' :render_width 32 :render_height 32 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 8)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,4, 0,0,0,1])
 (emission :name 'em' :type 'standard' :radiance (illuminant "D65"))
 (material :name 'm_floor' :type 'orennayar' :albedo (refl 0.7 0.7 0.6) :roughness 0.8)
 (material :name 'm_mirror' :type 'mirror')
 (material :name 'm_mirror_tint' :type 'reflection' :specularity (refl 0.9 0.6 0.2))
 (material :name 'm_oren' :type 'orennayar' :albedo (refl 0.2 0.5 0.8) :roughness 0.5)
 (material :name 'm_oren0' :type 'rough' :albedo (refl 0.8 0.3 0.3) :roughness 0)
 (material :name 'm_lamp' :type 'diffuse' :albedo 0.5)
 (entity :name 'floor' :type 'plane' :centering true :width 4 :height 4 :material 'm_floor' :position [0,0,-1])
 (entity :name 's0' :type 'sphere' :radius 0.5 :material 'm_mirror' :position [-0.9,0.7,0])
 (entity :name 's1' :type 'sphere' :radius 0.5 :material 'm_mirror_tint' :position [0.9,0.7,0])
 (entity :name 's2' :type 'sphere' :radius 0.5 :material 'm_oren' :position [-0.9,-0.7,0])
 (entity :name 's3' :type 'sphere' :radius 0.5 :material 'm_oren0' :position [0.9,-0.7,0])
 (entity :name 'lamp' :type 'plane' :centering true :width 1 :height 1 :material 'm_lamp' :emission 'em' :transform [1,0,0,1.2, 0,-1,0,1.2, 0,0,-1,3, 0,0,0,1])
 (light :type 'env' :radiance (illuminant "D65"))
)
"""

# blend.cpp / add.cpp: every MaterialDelta variant -- None, First, Second, All -- over leaf materials
MATERIAL_ZOO3 = """
(scene :name 'zoo3' :render_width 32 :render_height 32 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 8)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.1 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,4, 0,0,0,1])
 (emission :name 'em' :type 'standard' :radiance (illuminant "D65"))
 (material :name 'm_diffuse' :type 'diffuse' :albedo (refl 0.8 0.4 0.2))
 (material :name 'm_metal' :type 'roughconductor' :eta 0.2 :k 3.0 :roughness 0.2)
 (material :name 'm_mirror' :type 'mirror' :specularity (refl 0.9 0.9 0.8))
 (material :name 'm_glass' :type 'glass' :index 1.55)
 (material :name 'm_oren' :type 'orennayar' :albedo (refl 0.3 0.6 0.4) :roughness 0.6)
 (material :name 'b_none' :type 'blend' :material1 'm_diffuse' :material2 'm_metal' :factor 0.3)
 (material :name 'b_first' :type 'mix' :material1 'm_mirror' :material2 'm_diffuse' :factor 0.6)
 (material :name 'b_second' :type 'blend' :material1 'm_oren' :material2 'm_glass' :factor 0.4)
 (material :name 'b_all' :type 'blend' :material1 'm_mirror' :material2 'm_glass')
 (material :name 'a_none' :type 'add' :material1 'm_diffuse' :material2 'm_metal')
 (material :name 'a_first' :type 'add' :material1 'm_mirror' :material2 'm_oren')
 (entity :name 'floor' :type 'plane' :centering true :width 4 :height 4 :material 'm_diffuse' :position [0,0,-1])
 (entity :name 's0' :type 'sphere' :radius 0.4 :material 'b_none' :position [-1.0,0.7,0])
 (entity :name 's1' :type 'sphere' :radius 0.4 :material 'b_first' :position [0,0.7,0])
 (entity :name 's2' :type 'sphere' :radius 0.4 :material 'b_second' :position [1.0,0.7,0])
 (entity :name 's3' :type 'sphere' :radius 0.4 :material 'b_all' :position [-1.0,-0.5,0])
 (entity :name 's4' :type 'sphere' :radius 0.4 :material 'a_none' :position [0,-0.5,0])
 (entity :name 's5' :type 'sphere' :radius 0.4 :material 'a_first' :position [1.0,-0.5,0])
 (entity :name 'lamp' :type 'plane' :centering true :width 1 :height 1 :material 'm_diffuse' :emission 'em' :transform [1,0,0,1.2, 0,-1,0,1.2, 0,0,-1,3, 0,0,0,1])
 (light :type 'env' :radiance (illuminant "D65"))
)
"""

# The two scenes of the reference's end-to-end furnace test, src/tests/python/whitefurnance.py:10-140, verbatim (test
# fixtures: unit sphere, albedo 1, constant environment; orthographic camera, hammersley 8 spp, block filter radius 0).
# Use with .format(size=...) / .format(hero="true"|"false", size=...).
WHITEFURNACE_SPEC = """
(scene
    :name 'spectral_test'
    :render_width {size}
    :render_height {size}
    :spectral_domain 520

    ; Settings
    (integrator
        :type 'DIRECT'
        :max_ray_depth 4
        :light_sampe_count 1
    )
    (sampler
        :slot 'aa'
        :type 'hammersley'
        :sample_count 8
    )
    (filter
        :slot 'pixel'
        :type 'BLOCK'
        :radius 0
    )
    ; Outputs
    (output
        :name 'image'
        (channel :type 'color' :color 'srgb' )
    )
    ; Camera
    (camera
        :name 'Camera'
        :type 'orthographic'
        :width 2
        :height 2
        :local_direction [0,0,1]
        :local_up [0,1,0]
        :local_right [1,0,0]
        :position [0,0,-1.0005]
    )
    ; Background
    (light
        :name 'background'
        :type 'env'
        :radiance 1
    )
    ; Materials
    (material
        :name 'Diffuse'
        :type 'diffuse'
        :albedo 1
    )
    ; Primitives
    (entity
        :type "sphere"
        :name "Unit Sphere"
        :radius 1
        :material "Diffuse"
    )
)
"""

WHITEFURNACE_FULL = """
(scene
    :name 'illum_test'
    :render_width {size}
    :render_height {size}
    :camera 'Camera'
    :spectral_hero {hero}

    ; Settings
    (integrator
        :type 'DIRECT'
        :max_ray_depth 4
        :light_sampe_count 1
    )
    (sampler
        :slot 'aa'
        :type 'hammersley'
        :sample_count 8
    )
    (sampler
        :slot 'spectral'
        :type 'random'
        :sample_count 1
    )
    (filter
        :slot 'pixel'
        :type 'BLOCK'
        :radius 0
    )
    ; Outputs
    (output
        :name 'image'
        (channel :type 'color' :color 'srgb' )
    )
    ; Camera
    (camera
        :name 'Camera'
        :type 'orthographic'
        :width 2
        :height 2
        :local_direction [0,0,1]
        :local_up [0,1,0]
        :local_right [1,0,0]
        :position [0,0,-1.00005]
    )
    ; Background
    (light
        :name 'background'
        :type 'env'
        :radiance (illuminant "D65")
    )
    ; Materials
    (material
        :name 'Diffuse'
        :type 'diffuse'
        :albedo "white"
    )
    ; Primitives
    (entity
        :type "sphere"
        :name "Unit Sphere"
        :radius 1
        :material "Diffuse"
    )
)
"""


# light path expression channels: the material zoo (every scattering type; area light + environment) with one
# spectral channel per expression.  'C.*L' accepts every path, so its channel must equal the colour channel bit for bit.
LPE_EXPRESSIONS = ["C.*L", "CL", "C.+L", "C.*B", "C.*E", "CDL", "C<T,S>+.*L", "C<R,S>[DS]*E"]
LPE_ZOO = MATERIAL_ZOO.replace("(light :type 'env'", "(output :name 'image' (channel :type 'color' :color 'xyz')\n"
                               + "".join("   (channel :type 'color' :color 'xyz' :lpe '%s')\n" % e for e in LPE_EXPRESSIONS)
                               + "   (channel :type 'depth' :lpe 'C') (channel :type 'n' :lpe 'CD'))\n (light :type 'env'")


# nested blend / add materials (blend.cpp / add.cpp take any IMaterial as child): two and three levels, a delta
# subtree as first child, an all-delta tree
NESTED_MATERIALS = """ (material :name 'n2' :type 'blend' :material1 'b_none' :material2 'm_oren' :factor 0.5)
 (material :name 'n3' :type 'add' :material1 'n2' :material2 'a_first')
 (material :name 'n_delta_first' :type 'blend' :material1 'b_all' :material2 'n2' :factor 0.7)
 (material :name 'n_all' :type 'blend' :material1 'b_all' :material2 'm_mirror' :factor 0.25)
"""
MATERIAL_ZOO4 = (MATERIAL_ZOO3.replace(" (entity :name 'floor'", NESTED_MATERIALS + " (entity :name 'floor'")
                 .replace(":material 'b_none' :position", ":material 'n2' :position").replace(":material 'b_first' :position", ":material 'n3' :position")
                 .replace(":material 'b_second' :position", ":material 'n_delta_first' :position").replace(":material 'b_all' :position", ":material 'n_all' :position"))


# depth of field (PerspectiveCamera<HasDOF>, perspective.cpp:66-75): the material zoo through a lens; {dof} = camera options
DOF_ZOO = MATERIAL_ZOO.replace(":near 0.1 :far 100", ":near 0.1 :far 100 {dof}")
