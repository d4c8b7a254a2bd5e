--  LZ4Ada.Device -- thin binding of the device shim (include/lz4b200.h).
--  UNCOMPILED in this image (no GNAT); see ada/README.md.
--  Every function returns the shim's int: 0 ok, negative = CUDA / argument failure.  LZ4 data
--  errors never travel in the return code; they come back in Block_Status records.
with Interfaces;   use Interfaces;
with Interfaces.C; use Interfaces.C;
with Interfaces.C.Strings;
with System;

private package LZ4Ada.Device is

   type Context is new System.Address;   --  lz4b200_ctx *
   type Stream  is new System.Address;   --  lz4b200_stream *

   --  lz4b200_blk_desc
   type Block_Desc is record
      Src_Off    : Unsigned_64;   --  first payload byte (after the 4-byte size word)
      Dst_Off    : Unsigned_64;   --  where the block's output starts
      Src_Len    : Unsigned_32;   --  payload bytes, without the checksum trailer
      Dst_Cap    : Unsigned_32;
      Flags      : Unsigned_32;   --  Blk_Stored or Blk_Has_Checksum or ...
      Hist_Avail : Unsigned_32;   --  frame position of the block start
   end record with Convention => C;

   Blk_Stored         : constant Unsigned_32 := 1;
   Blk_Has_Checksum   : constant Unsigned_32 := 2;
   Blk_Hash_Only      : constant Unsigned_32 := 4;
   Blk_Chained        : constant Unsigned_32 := 8;
   Blk_First_Of_Frame : constant Unsigned_32 := 16;
   Blk_Solo           : constant Unsigned_32 := 32;    --  chain of one block taken from an independent frame
   Blk_Ring_Cap       : constant Unsigned_32 := 64;    --  Dst_Cap = the caller's Buffer length (lib/lz4ada.adb:54, 678-680)
   Blk_K2             : constant Unsigned_32 := 128;   --  stored block routed to Copy_Stored; K1 skips it

   --  lz4b200_blk_status
   type Block_Status is record
      Code           : Unsigned_32;   --  0 ok, 1 block checksum, 2 ends after literals, 3 offset 0,
                                      --  4 back-reference range, 5..9 Appendix C cases,
                                      --  10 needs history (soft), 11 not run
      Out_Len        : Unsigned_32;
      Err_Pos        : Unsigned_32;
      Aux            : Integer_32;
      XXH32_Computed : Unsigned_32;
      XXH32_Declared : Unsigned_32;
   end record with Convention => C;

   type Chain is record
      First_Block, N_Blocks : Unsigned_32;
      Dst_Off, Dst_Cap      : Unsigned_64;
   end record with Convention => C;

   type Frame_Blocks is record
      First_Block, N_Blocks : Unsigned_32;
   end record with Convention => C;

   function Create (Device_Index : int; Cuda_Stream : System.Address; Ctx : out Context) return int
     with Import, Convention => C, External_Name => "lz4b200_create";
   --  A context is reference counted: Create returns one reference, Destroy drops one, Retain takes one more (a
   --  Limited_Controlled Decompressor holds its own, so finalisation order does not matter).
   function Retain (Ctx : Context) return int
     with Import, Convention => C, External_Name => "lz4b200_retain";
   function Destroy (Ctx : Context) return int
     with Import, Convention => C, External_Name => "lz4b200_destroy";
   function Last_Error (Ctx : Context) return Interfaces.C.Strings.chars_ptr
     with Import, Convention => C, External_Name => "lz4b200_last_error";

   function Alloc (Ctx : Context; Bytes : size_t; Ptr : out System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_alloc";
   function Free (Ctx : Context; Ptr : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_free";
   function Alloc_Host (Ctx : Context; Bytes : size_t; Ptr : out System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_alloc_host";
   function Free_Host (Ctx : Context; Ptr : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_free_host";
   function H2D (Ctx : Context; Dst_Dev, Src_Host : System.Address; Bytes : size_t) return int
     with Import, Convention => C, External_Name => "lz4b200_h2d";
   function D2H (Ctx : Context; Dst_Host, Src_Dev : System.Address; Bytes : size_t) return int
     with Import, Convention => C, External_Name => "lz4b200_d2h";
   function Sync (Ctx : Context) return int
     with Import, Convention => C, External_Name => "lz4b200_sync";

   --  K1 (+K2): independent blocks, fused block XXH32   (lib/lz4ada.adb:661-904)
   function Decode_Blocks (Ctx : Context; Src, Dst : System.Address; N_Blocks : Unsigned_32;
                           Desc, Status : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_decode_blocks";
   --  K4: chains (linked frames, exact-placement retries)
   function Decode_Linked (Ctx : Context; Src, Dst : System.Address; N_Chains : Unsigned_32;
                           Chains, Desc, Status : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_decode_linked";
   --  K3: content checksum per frame                      (lib/lz4ada.adb:709-714, 942-1017)
   function XXH32_Frames (Ctx : Context; Dst : System.Address; N_Frames : Unsigned_32;
                          Frames, Desc, Status, Digest, Valid : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_xxh32_frames";
   --  K3 over explicit byte ranges (content checksums of re-placed frames, block checksums of stored blocks)
   type Hash_Span is record
      Off, Len : Unsigned_64;
   end record with Convention => C;
   function XXH32_Spans (Ctx : Context; Data : System.Address; N : Unsigned_32; Spans, Digests : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_xxh32_spans";
   --  K2: stored blocks as a wide copy (lib/lz4ada.adb:685-695), block checksums beside it
   function Copy_Stored (Ctx : Context; Src, Dst : System.Address; N_Idx : Unsigned_32; Idx : System.Address;
                         Max_Len : Unsigned_32; Desc, Status, Spans, Scratch : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_copy_stored";
   --  K5: size pre-pass
   function Size_Blocks (Ctx : Context; Src : System.Address; N_Blocks : Unsigned_32;
                         Desc, Status : System.Address) return int
     with Import, Convention => C, External_Name => "lz4b200_size_blocks";

   --  single-block path under Decompressor.Update
   function Stream_Create (Ctx : Context; Max_Block : Unsigned_32; S : out Stream) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_create";
   function Stream_Destroy (S : Stream) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_destroy";
   function Stream_Reset (S : Stream) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_reset";
   function Stream_Block (S : Stream; Host_Src : System.Address; Src_Len, Flags : Unsigned_32;
                          Hash_Content : int; Host_Dst : System.Address; Dst_Cap : Unsigned_32;
                          Status : out Block_Status) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_block";
   --  ... with the frame's block maximum as a hint: one synchronisation per block instead of two
   function Stream_Block2 (S : Stream; Host_Src : System.Address; Src_Len, Flags : Unsigned_32;
                           Hash_Content : int; Host_Dst : System.Address; Dst_Cap, Expect_Out : Unsigned_32;
                           Status : out Block_Status) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_block2";
   function Stream_Digest (S : Stream; XXH32 : out Unsigned_32) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_digest";
   --  read-ahead under Update: N bytes decoded ahead of time by Decode_Blocks into a device staging buffer
   --  become the stream's next bytes (history window + running content checksum)
   function Stream_Adopt (S : Stream; Dev_Bytes : System.Address; N : Unsigned_32; Hash_Content : int) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_adopt";
   --  ... up to 255 served blocks at once (Offsets / Lengths: arrays of Unsigned_32, pieces in stream order)
   function Stream_Adopt_List (S : Stream; Dev_Base : System.Address; N_Pieces : Unsigned_32;
                               Offsets, Lengths : System.Address; Hash_Content : int) return int
     with Import, Convention => C, External_Name => "lz4b200_stream_adopt_list";

   --  which K1 kernel Decode_Blocks launches: 0 = chosen from the block count (v6 lane-per-block from ~20 000
   --  blocks on, v4 warp-per-block below); see include/lz4b200.h for the other values
   function Set_Tuning (Ctx : Context; Generation : int) return int
     with Import, Convention => C, External_Name => "lz4b200_set_tuning";
   function K1_Kernel_Name (Ctx : Context; N_Blocks : Unsigned_32) return Interfaces.C.Strings.chars_ptr
     with Import, Convention => C, External_Name => "lz4b200_k1_kernel_name";

   --  statistics for a maintainer: blocks the lane-per-block K1 (v6) handed to the exact routine in its last launch,
   --  blocks the chain kernel (K7) finished / gave up on since the last call
   function K1_Fallbacks (Ctx : Context; To_Exact, By_Safety_Net : out Unsigned_32) return int
     with Import, Convention => C, External_Name => "lz4b200_k1_fallbacks";
   function Chain_Stats (Ctx : Context; Finished, Given_Up : out Unsigned_32) return int
     with Import, Convention => C, External_Name => "lz4b200_chain_stats";

end LZ4Ada.Device;
