// OceanNative.cs -- P/Invoke binding of libocean.so (include/ocean.h) and a WaveGenerator-shaped host class.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no dotnet/mono/csc.  The ABI is exercised through
// the Python ctypes mirror (godotoceanwaves_b200/native.py) and tests/test_abi_cpu.py instead; this file shows the
// binding a Godot C# project would add.  Convention follows the one in-tree P/Invoke precedent of the reference,
// addons/imgui-godot/ImGuiGodot/Internal/Viewports.cs:157-165 ([LibraryImport], cdecl, unsafe partial).
//
// Replaces: assets/water/wave_generator.gd (WaveGenerator node) and, below it, every RenderingContext call of
// assets/render_context.gd:35-135 that the generator makes.  The two RGBA16F layered maps are handed to Godot with
// RenderingDevice.TextureUpdate(rid, layer, bytes), which the reference's textures already allow
// (TEXTURE_USAGE_CAN_UPDATE_BIT, wave_generator.gd:34-35).
using System;
using System.Runtime.InteropServices;

namespace OceanB200
{
    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct OceanCascadeParams            // struct ocean_cascade_params <- wave_cascade_parameters.gd:2-42
    {
        public fixed float tile_length[2];
        public double displacement_scale, normal_scale;
        public double wind_speed, wind_direction, fetch_length, swell, spread, detail, whitecap, foam_amount;
        public fixed int spectrum_seed[2];
        public int should_generate_spectrum;
        public double time, foam_grow_rate, foam_decay_rate;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct OceanInfo
    {
        public int device, map_size, num_cascades, pending_cascades;
        public ulong kernel_launches, cascade_updates, device_bytes;
    }

    internal static unsafe partial class Native
    {
        private const string Lib = "ocean";           // libocean.so next to the Godot binary / in LD_LIBRARY_PATH

        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_default_cascade_params(OceanCascadeParams* p);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_create(int device, int map_size, int num_cascades, IntPtr* handle);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_destroy(IntPtr handle);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_update(IntPtr handle, double delta, OceanCascadeParams* parameters, int count);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_process(IntPtr handle, OceanCascadeParams* parameters, int count);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_update_all(IntPtr handle, double delta, OceanCascadeParams* parameters, int count);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_get_maps(IntPtr handle, IntPtr* displacement_dev, IntPtr* normal_dev, nuint* layer_bytes);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_copy_maps_to_host(IntPtr handle, int first, int count, void* displacement, void* normal);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_copy_maps_to_host_async(IntPtr handle, int first, int count, void* displacement, void* normal);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_synchronize(IntPtr handle);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_host_alloc(void** ptr, nuint bytes);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_host_free(void* ptr);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_get_foam_state(IntPtr handle, int cascade, ushort* host);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_set_foam_state(IntPtr handle, int cascade, ushort* host);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial double ocean_jonswap_alpha(double wind_speed, double fetch_length);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial double ocean_jonswap_peak_angular_frequency(double wind_speed, double fetch_length);
        // map queries (water.gdshader:27-39,42-84): points [n][2] world x,z; map_scales [c][4]; outputs [n][3]
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_sample_maps(IntPtr handle, int num_points, float* points_xz, int num_cascades, float* map_scales,
                                                      float* displacement, float* gradient_foam);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial int ocean_get_info(IntPtr handle, OceanInfo* info);
        [LibraryImport(Lib)] [UnmanagedCallConv(CallConvs = new[] { typeof(System.Runtime.CompilerServices.CallConvCdecl) })]
        internal static partial IntPtr ocean_last_error();
    }

    public sealed class OceanException : Exception { public OceanException(string m) : base(m) { } }

    /// <summary>Same surface as the reference's WaveGenerator (assets/water/wave_generator.gd): MapSize, InitGpu,
    /// Update, Process, the two maps and the static JONSWAP helpers.</summary>
    public sealed unsafe class WaveGenerator : IDisposable
    {
        public const double G = 9.81, Depth = 20.0;               // wave_generator.gd:5-6
        public int MapSize;                                         // wave_generator.gd:8
        private IntPtr _handle;
        private int _layers;
        private byte* _dispHost, _normHost;                         // pinned staging for texture_update

        private static void Check(int status)
        {
            if (status != 0) throw new OceanException(Marshal.PtrToStringAnsi(Native.ocean_last_error()) ?? "libocean error");
        }

        public void InitGpu(int numCascades, int device = 0)        // wave_generator.gd:17-54
        {
            Dispose();
            IntPtr h;
            Check(Native.ocean_create(device, MapSize, numCascades, &h));
            _handle = h; _layers = numCascades;
            nuint bytes = (nuint)numCascades * (nuint)MapSize * (nuint)MapSize * 8;
            void* d, n;
            Check(Native.ocean_host_alloc(&d, bytes)); Check(Native.ocean_host_alloc(&n, bytes));
            _dispHost = (byte*)d; _normHost = (byte*)n;
        }

        public void Update(double delta, Span<OceanCascadeParams> parameters)   // wave_generator.gd:90-109
        {
            if (parameters.Length == 0) throw new ArgumentException("parameters.size() != 0");
            if (_handle == IntPtr.Zero) InitGpu(Math.Max(2, parameters.Length));
            fixed (OceanCascadeParams* p = parameters) Check(Native.ocean_update(_handle, delta, p, parameters.Length));
        }

        public void Process(Span<OceanCascadeParams> parameters)                 // wave_generator.gd:56-63 (_process)
        {
            fixed (OceanCascadeParams* p = parameters) Check(Native.ocean_process(_handle, p, parameters.Length));
        }

        /// <summary>Copies layer `cascade` of both maps to pinned host memory; the caller passes the spans to
        /// RenderingDevice.TextureUpdate(displacementRid, cascade, bytes) / (normalRid, cascade, bytes).</summary>
        public (IntPtr displacement, IntPtr normal, int bytes) FetchLayer(int cascade)
        {
            int layer = MapSize * MapSize * 8;
            Check(Native.ocean_copy_maps_to_host(_handle, cascade, 1, _dispHost + (long)cascade * layer, _normHost + (long)cascade * layer));
            return ((IntPtr)(_dispHost + (long)cascade * layer), (IntPtr)(_normHost + (long)cascade * layer), layer);
        }

        public static double JONSWAPAlpha(double windSpeed = 20.0, double fetchLength = 550e3) => Native.ocean_jonswap_alpha(windSpeed, fetchLength);
        public static double JONSWAPPeakAngularFrequency(double windSpeed = 20.0, double fetchLength = 550e3) => Native.ocean_jonswap_peak_angular_frequency(windSpeed, fetchLength);

        public void Dispose()                                       // NOTIFICATION_PREDELETE, wave_generator.gd:111-113
        {
            if (_handle != IntPtr.Zero) { Native.ocean_destroy(_handle); _handle = IntPtr.Zero; }
            if (_dispHost != null) { Native.ocean_host_free(_dispHost); _dispHost = null; }
            if (_normHost != null) { Native.ocean_host_free(_normHost); _normHost = null; }
        }
    }
}
