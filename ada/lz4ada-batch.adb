--  LZ4Ada.Batch body -- UNCOMPILED sketch (no GNAT in this image).  Statement-level twin of
--  bo_lz4_ada_b200/csrc/host/batch.cpp; that file is the one the tests exercise.
--
--  Stages
--    1. Plan (host): walk every stream with the package's own header parser and size-word
--       reader (Process_Header_Bytes lib/lz4ada.adb:155, Try_Detect_Input_Length :525) but, instead
--       of decoding a complete block, append a Device.Block_Desc {Src_Off, Src_Len, Stored,
--       Has_Checksum}.  Header / size-word / Single_Frame exceptions are caught per stream and kept
--       as that stream's host outcome -- they are, by construction, later in stream order than every
--       recorded block.
--    2. Place: block i of a frame goes to Frame_Base + i * Block_Max (the frame format carries no
--       per-block decompressed size; every mainstream encoder fills its blocks).  The last block of
--       a frame that is followed by another frame of the same stream is sized first by K5.
--    3. Run (device): Decode_Blocks over all blocks (chained ones are skipped by the kernel),
--       Decode_Linked over the linked frames, XXH32_Frames over the frames with a content checksum;
--       one D2H of the status / digest arrays.
--    4. Fold (host): in stream order, per block: content-size overflow (Decrease_Data_Size_Remaining
--       :826) using Out_Len / Err_Pos, then the block's own status -> the reference's exception and
--       message; per frame: content checksum (:505), content size left (:471).  First one wins;
--       otherwise the host outcome of stage 1.
--    5. Streams whose placement assumption broke (short interior block, a match into the previous
--       block of an "independent" frame) are decoded again as one chain with exact running
--       placement (Decode_Linked) and folded again.
with LZ4Ada.Device;

package body LZ4Ada.Batch is

   function Output_Bytes (Source : in Octets; Spans : in Stream_Spans;
                          Reservation : in Memory_Reservation := For_All) return U64 is
   begin
      raise Program_Error with "LZ4Ada.Batch: Ada body not built in this image; see host/batch.cpp";
      return 0;
   end Output_Bytes;

   procedure Decompress (Source      : in     Octets;
                         Spans       : in out Stream_Spans;
                         Destination :    out Octets;
                         Results     :    out Outcomes;
                         Reservation : in     Memory_Reservation := For_All;
                         GPU         : in     Natural := 0) is
   begin
      raise Program_Error with "LZ4Ada.Batch: Ada body not built in this image; see host/batch.cpp";
   end Decompress;

end LZ4Ada.Batch;
